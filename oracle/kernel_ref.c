/*
 * Plain-C restatement of the reference CUDA kernels' arithmetic.
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from
 * the product (boxer_b200).  Used by tests/ as a second, independent checker
 * (it states the *kernel* semantics, where oracle/plain.py states the
 * grid_sample semantics) and pinned against the tests/golden npz fixtures.
 *
 * Follows, without copying:
 *   /root/reference/e2edet/module/ops/src/box_attn/box_attn_kernel.cuh
 *       :34-97    bilinear sample with per-corner zero padding
 *       :100-184  backward of one sample (grad_value scatter, grad_loc, grad_attn)
 *       :274-349  forward loop (pixel = loc*size - 0.5, window test :328)
 *   /root/reference/e2edet/module/ops/src/instance_attn/instance_attn_kernel.cuh
 *       :98-187   backward of one sample with two weights (:139, :183-186)
 *       :282-364  forward with the per-point mask output (:354-355)
 *
 * Layouts (all contiguous, reference order):
 *   value      (B, S, H, D)            shapes (L, 2) int64 = (h_l, w_l)
 *   loc        (B, Nq, H, L, P, 2)     (x, y) normalised to [0, 1]
 *   weights    (B, Nq, H, L, P)
 *   out        (B, Nq, H, D)
 *   mask_out   (B, Nq, P, H, D)        (instance op only)
 *
 * Built twice from one body: REAL = double and REAL = float.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL double
#define SUF _f64
#include "kernel_ref_body.inc"
#undef REAL
#undef SUF

#define REAL float
#define SUF _f32
#include "kernel_ref_body.inc"
#undef REAL
#undef SUF

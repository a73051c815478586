"""compat.install(): the reference's import paths must end up bound to this package's classes -- with a real
checkout on sys.path (install before or after ``import e2edet``), with a checkout whose package __init__ cannot
run, and with no checkout at all.  (the first version imported ``e2edet`` itself inside install(), so
the reference classes were bound before the override, or an import error was swallowed into an empty placeholder.)"""
import os
import sys
import textwrap
import types

import pytest


def _purge():
    for n in [n for n in sys.modules if n == "e2edet" or n.startswith("e2edet.")]:
        sys.modules.pop(n, None)


@pytest.fixture
def clean_e2edet():
    import boxer_b200
    boxer_b200.compat.uninstall()
    _purge()
    yield
    boxer_b200.compat.uninstall()
    _purge()


def _fake_checkout(root, init_body="import e2edet.module\n", pybind_body="raise ImportError('e2edet.ops: extension not built')\n"):
    """A miniature of the reference's package layout and import statements (e2edet/__init__.py,
    module/__init__.py, module/box_transformer.py:6, module/ops/__init__.py:1-3, box_attention_func.py:6)."""
    files = {
        "e2edet/__init__.py": init_body,
        "e2edet/ops.py": pybind_body,
        "e2edet/utils/__init__.py": "",
        "e2edet/utils/general.py": "def get_clones(m, n):\n    return [m] * n\n",
        "e2edet/module/__init__.py": "from e2edet.module.transformer import build_transformer\n",
        "e2edet/module/transformer.py": "from e2edet.module.box_transformer import BoxTransformer\n"
                                        "from e2edet.module.box3d_transformer import Box3dTransformer\n"
                                        "def build_transformer():\n    return BoxTransformer()\n",
        "e2edet/module/box_attention.py": "from e2edet.module.ops import BoxAttnFunction, InstanceAttnFunction\n"
                                          "class BoxAttention:\n    REFERENCE = True\n"
                                          "class InstanceAttention:\n    REFERENCE = True\n"
                                          "class Box3dAttention:\n    REFERENCE = True\n",
        "e2edet/module/ops/__init__.py": "from .box_attention_func import BoxAttnFunction, InstanceAttnFunction\n",
        "e2edet/module/ops/box_attention_func.py": "from e2edet import ops\n"
                                                   "class BoxAttnFunction:\n    REFERENCE = True\n"
                                                   "class InstanceAttnFunction:\n    REFERENCE = True\n",
        "e2edet/module/box_transformer.py": textwrap.dedent("""
            from .box_attention import BoxAttention, InstanceAttention
            from e2edet.utils.general import get_clones
            class BoxTransformer:
                def __init__(self):
                    self.self_attn = BoxAttention(32, 2, 4)
                    self.mask_attn = InstanceAttention(32, 2, 4, 4)
            """),
        "e2edet/module/box3d_transformer.py": "from .box_attention import Box3dAttention\n"
                                              "class Box3dTransformer:\n"
                                              "    def __init__(self):\n        self.attn = Box3dAttention(32, 2, 4)\n",
    }
    for rel, body in files.items():
        path = os.path.join(root, rel)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(body)


def test_install_before_import_binds_our_classes(tmp_path, monkeypatch, clean_e2edet):
    import boxer_b200
    _fake_checkout(str(tmp_path))
    monkeypatch.syspath_prepend(str(tmp_path))
    boxer_b200.compat.install()
    assert "e2edet" not in sys.modules                      # install() does not run the checkout's package code
    import e2edet                                           # the real __init__ chain runs now
    from e2edet.module import box_transformer as bt, box3d_transformer as b3
    assert e2edet.__file__.startswith(str(tmp_path))
    assert bt.BoxAttention is boxer_b200.BoxAttention and bt.InstanceAttention is boxer_b200.InstanceAttention
    assert b3.Box3dAttention is boxer_b200.Box3dAttention
    layer = bt.BoxTransformer()
    assert isinstance(layer.self_attn, boxer_b200.BoxAttention) and isinstance(layer.mask_attn, boxer_b200.InstanceAttention)
    from e2edet.module.ops import BoxAttnFunction
    from e2edet import ops
    assert BoxAttnFunction is boxer_b200.BoxAttnFunction and ops is boxer_b200.ops      # the pybind stub was never loaded


def test_install_after_import_rebinds_loaded_modules(tmp_path, monkeypatch, clean_e2edet):
    import boxer_b200
    _fake_checkout(str(tmp_path), pybind_body="def box_attn_forward(*a):\n    raise RuntimeError('reference pybind')\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    import e2edet
    from e2edet.module import box_transformer as bt
    from e2edet.module.ops import box_attention_func as ref_func
    assert getattr(bt.BoxAttention, "REFERENCE", False)
    boxer_b200.compat.install()
    assert bt.BoxAttention is boxer_b200.BoxAttention and bt.InstanceAttention is boxer_b200.InstanceAttention
    assert sys.modules["e2edet.module.box3d_transformer"].Box3dAttention is boxer_b200.Box3dAttention
    assert ref_func.ops is boxer_b200.ops                   # Functions someone still holds call this library
    assert isinstance(bt.BoxTransformer().self_attn, boxer_b200.BoxAttention)
    from e2edet.module.box_attention import BoxAttention
    assert BoxAttention is boxer_b200.BoxAttention
    boxer_b200.compat.uninstall()
    assert getattr(bt.BoxAttention, "REFERENCE", False)     # and back


def test_checkout_whose_init_fails_is_not_swallowed(tmp_path, monkeypatch, clean_e2edet):
    import boxer_b200
    _fake_checkout(str(tmp_path), init_body="import omegaconf_that_is_not_installed\n")
    monkeypatch.syspath_prepend(str(tmp_path))
    boxer_b200.compat.install()
    with pytest.raises(ModuleNotFoundError, match="omegaconf_that_is_not_installed"):
        import e2edet  # noqa: F401   (the checkout's own problem surfaces; no empty placeholder hides the package)
    boxer_b200.compat.uninstall()
    _purge()
    boxer_b200.compat.install(lightweight=True)             # bare parents with the checkout's real __path__
    from e2edet.module.box_transformer import BoxTransformer
    assert isinstance(BoxTransformer().self_attn, boxer_b200.BoxAttention)
    assert sys.modules["e2edet.utils.general"].__file__.startswith(str(tmp_path))


def test_no_checkout_placeholders(clean_e2edet):
    import boxer_b200
    boxer_b200.compat.install()
    from e2edet.module.ops import BoxAttnFunction
    from e2edet.module.box_attention import Box3dAttention
    from e2edet import ops
    assert BoxAttnFunction is boxer_b200.BoxAttnFunction and Box3dAttention is boxer_b200.Box3dAttention
    assert callable(ops.box_attn_forward)
    boxer_b200.compat.uninstall()
    assert "e2edet" not in sys.modules


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (authoring container only)")
def test_reference_checkout_layers_use_our_modules(monkeypatch, clean_e2edet):
    """The reference's own box_transformer.py / box3d_transformer.py, imported unmodified from the checkout
    (its package __init__ needs omegaconf / pycocotools / torch._six, so: lightweight parents + the one-line
    torch._six shim general.py:12 needs on torch >= 2), build their layers from this package's classes."""
    import boxer_b200
    monkeypatch.syspath_prepend(REF)
    six = types.ModuleType("torch._six")
    six.string_classes = (str, bytes)
    monkeypatch.setitem(sys.modules, "torch._six", six)
    boxer_b200.compat.install(lightweight=True)
    from e2edet.module.box_transformer import BoxTransformer, BoxTransformerEncoderLayer
    from e2edet.module.box3d_transformer import Box3dTransformer
    assert sys.modules["e2edet.module.box_transformer"].__file__.startswith(REF)
    t = BoxTransformer(d_model=64, nhead=4, nlevel=2, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64,
                       num_queries=5, use_mask=True)
    assert isinstance(t.encoder.layers[0].self_attn, boxer_b200.BoxAttention)
    assert isinstance(t.decoder.layers[0].multihead_attn, boxer_b200.InstanceAttention)
    t3 = Box3dTransformer(d_model=64, nhead=4, nlevel=2, num_encoder_layers=1, num_decoder_layers=1, dim_feedforward=64, num_queries=5)
    assert isinstance(t3.encoder.layers[0].self_attn, boxer_b200.Box3dAttention)
    assert isinstance(t3.decoder.layers[0].multihead_attn, boxer_b200.Box3dAttention)
    assert isinstance(BoxTransformerEncoderLayer(64, 4, 2, 64, 0.0, "relu").self_attn, boxer_b200.BoxAttention)

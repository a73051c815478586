"""The reference's own caller code around the hot path (compiled, never copied): see build_ref_layers.py."""

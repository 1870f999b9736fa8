"""Harnesses around the warp for BASELINE configs 3 and 4 (training step, clip inference): the callers of the hot
path replayed on synthetic data.  Not part of the product package; bench.py --config train|clip runs them."""

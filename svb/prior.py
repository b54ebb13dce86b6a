"""Placeholder namespace: aslnn.py:29 imports svb.prior without using it; priors live in the fused kernel."""

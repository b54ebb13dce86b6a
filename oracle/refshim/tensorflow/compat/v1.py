"""`import tensorflow.compat.v1 as tf` (aslrest.py:4-7) resolves to the same numpy stand-in."""
from tensorflow import *  # noqa: F401,F403
from tensorflow import math, nn, Session, Tensor, float32, float64  # noqa: F401

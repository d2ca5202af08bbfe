"""Drop-in shim: put this directory ahead of the reference on sys.path and the reference's
main.py / worker.py import the B200-native `ema` unchanged."""
from lcgan_b200.ema import *  # noqa: F401,F403
from lcgan_b200.ema import Ema  # noqa: F401

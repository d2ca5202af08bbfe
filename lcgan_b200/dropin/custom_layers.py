"""Drop-in shim: put this directory ahead of the reference on sys.path and the reference's
main.py / worker.py import the B200-native `custom_layers` unchanged."""
from lcgan_b200.custom_layers import *  # noqa: F401,F403

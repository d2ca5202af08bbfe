"""Import stub: PyAV is only used by the reference's mp4 writers (worker.py:255-379), not by training."""

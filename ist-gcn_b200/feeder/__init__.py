"""Drop-in for the reference ``feeder`` package (feeder/feeder.py, feeder/tools.py)."""

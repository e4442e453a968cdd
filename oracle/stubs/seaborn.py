"""Empty stand-in: /root/reference/utils.py:5 imports seaborn but the hot path never calls it."""

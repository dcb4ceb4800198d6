from .freq import FreqEncoder, freq_encode  # noqa: F401  (encoding.py:16 does `from freqencoder import FreqEncoder`)

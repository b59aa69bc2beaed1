"""Core of the B200-native Segment-Anything-NeRF render path (ctypes binding + ops)."""

"""B200-native PP-YOLO hot path: ctypes binding of the C-ABI kernel library + the inference engine."""
__version__ = '0.1.0'

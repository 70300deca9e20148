"""B200-native PP-YOLO module surface (drop-in for the reference's ``model`` package)."""

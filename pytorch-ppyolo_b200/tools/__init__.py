"""Host-side harness code mirroring the reference's ``tools`` package for the callers of the hot path."""

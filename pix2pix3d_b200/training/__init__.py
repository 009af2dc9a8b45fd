"""Host-side mirror of the reference's `training` package (generator side of the hot path)."""

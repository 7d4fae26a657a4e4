"""Host-side mirror of the reference's `utils` package for the hot path (models only)."""

"""Mirror of the reference's `models` package surface for the hot path (same symbol names)."""

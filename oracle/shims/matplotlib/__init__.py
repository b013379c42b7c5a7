"""matplotlib stub (oracle side only; the reference imports pyplot but never plots on the hot path)."""

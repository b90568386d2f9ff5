"""Data-side mirror of the reference's ``data_utils`` package (only what feeds the attack hot path)."""

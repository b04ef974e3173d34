"""Ensemble facades of the hot path (reference tlbo/facade/)."""

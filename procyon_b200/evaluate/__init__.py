"""Evaluation-framework plugins of the hot path (mirror of procyon/evaluate/framework/procyon.py)."""

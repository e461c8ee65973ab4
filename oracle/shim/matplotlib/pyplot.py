"""ORACLE ONLY: empty stand-in (reference: kernel/go_model.py:14)."""

"""ORACLE ONLY: empty stand-in (reference: util/image_cluster.py:9)."""

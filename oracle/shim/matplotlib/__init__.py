"""ORACLE ONLY: empty stand-in so the reference files import without plotting libs."""

"""Shadow package for the reference's `kernel` package: only the hot-path modules are provided."""

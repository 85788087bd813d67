"""Mirror of the reference's `utils/` package for the accelerated path."""

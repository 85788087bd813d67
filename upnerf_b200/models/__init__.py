"""Mirror of the reference's `models/` package for the accelerated path."""

"""diffbir.utils -> diffbir_b200.utils."""

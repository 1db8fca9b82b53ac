"""Host-side mirror of cupy/_core for the elementwise / reduction / scan hot path."""

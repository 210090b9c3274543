"""Empty stand-in: the reference imports matplotlib only for plotting helpers (TEST INFRASTRUCTURE)."""

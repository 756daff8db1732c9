"""Index directory I/O, construction and mutation (host side)."""

"""Version of the B200 build: the reference package version it mirrors (0.5) + a local build tag."""
__version__ = "0.5+b200.1"

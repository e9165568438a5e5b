__version__ = "0.5+b200.1"

"""nanoRLHF-B200: a Blackwell-native RLHF training engine (see DESIGN.md / SURVEY.md)."""
__version__ = "0.1.0"

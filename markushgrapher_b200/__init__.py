"""markushgrapher_b200 — B200-native hot path (image -> CXSMILES token ids) of MarkushGrapher-2."""
__version__ = "0.1.0"

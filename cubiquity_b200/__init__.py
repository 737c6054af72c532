"""cubiquity_b200 -- Blackwell-native SVDAG ray traversal and path tracing behind Cubiquity's raycast entry points."""
__version__ = "0.1.0"

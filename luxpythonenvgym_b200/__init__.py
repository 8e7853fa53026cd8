"""B200-native batched Lux AI 2021 turn engine (drop-in for the reference's per-turn path)."""
__version__ = "0.1.0"

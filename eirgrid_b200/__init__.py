"""eirgrid_b200 — B200-native batched 2025–2050 episode rollout for GridAI (ETM-Code/eirgrid)."""
__version__ = "0.1.0"

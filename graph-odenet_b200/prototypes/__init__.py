"""``prototypes/`` of the reference: only the Interaction-Network models of ``prototypes/orbit``."""

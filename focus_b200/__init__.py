"""focus_b200: B200-native (sm_100a) implementation of FOCUS's video slot-attention encoder."""
from .slot_attention import SlotAttentionVideo  # noqa: F401

__all__ = ["SlotAttentionVideo"]

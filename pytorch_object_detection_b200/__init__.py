"""B200-native FCOS post-processing / target assignment / loss hot path."""

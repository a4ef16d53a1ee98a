"""mtn_b200: B200-native (sm_100a) implementation of the MTN multimodal attention hot path."""

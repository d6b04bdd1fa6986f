"""graph_framework_b200: B200-native back end for graph_framework's per-ray hot path."""

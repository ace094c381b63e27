"""contourist_b200: B200-native engine behind the contourist class API (see README.md)."""

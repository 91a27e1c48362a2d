"""B200-native simplex pivot engine behind the `lpsol` API (placeholder, filled in below)."""

"""TEST INFRASTRUCTURE ONLY — import-time dummy (the reference touches pyplot even with render=False)."""

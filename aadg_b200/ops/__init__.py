"""Python faces of the C ABI entry points (torch tensors in, torch tensors out)."""

"""Host-side mirrors of the reference's Python interfaces around the hot path (controller, discriminator,
losses, config, the search step).  Small torch modules: scalar/latency work, not accelerated (SURVEY.md §2
rows 7-9, 11), kept so a run.py-style driver finds the call shapes it expects."""

"""Drop-in import paths of the reference (``src.env_definitions``, ``src.actions``, ``src.runs``,
``src.stats``, ``src.ppo``): put ``2048-ppo-agent_b200/`` on sys.path instead of the reference's
repo root and the same imports resolve to the B200 engine (package ``g2048``)."""

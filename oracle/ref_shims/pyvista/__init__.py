"""Empty stand-in: the reference imports pyvista at module level (rocket_env.py:8) but only
uses it for rendering, which is out of the hot path.  Test infrastructure only."""

"""Empty pyplot stand-in (see package docstring); nothing on the golden path plots."""

"""Import shim — TEST INFRASTRUCTURE ONLY: wrappers.py sets `pd.options.plotting.backend = "plotly"`,
which makes pandas import a module of that name exposing `plot`."""


def plot(*a, **k):
    raise RuntimeError("plotting is not available in the golden-vector generator")

def make_subplots(*a, **k):
    raise NotImplementedError("plotly stand-in")

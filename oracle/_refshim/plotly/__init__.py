"""Empty stand-in so that `import lightplane` (which hard-imports plotly in visualize.py:16-23)
works inside oracle/make_golden.py.  Nothing here is ever called."""

"""Inert stand-in for matplotlib (alpha_net.py:7-9 imports it; only train() uses it)."""


def use(*a, **k):
    return None

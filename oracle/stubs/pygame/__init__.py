"""Inert stand-in for pygame, used ONLY when the Python reference is imported
as a checker inside the build container (oracle/ref_harness.py).  The engine
path of the reference touches nothing but ``pg.Rect(...)`` at construction
time (inventory_frame.py:29-32,48-50, turn_panel.py:21-22)."""


class Rect:
    def __init__(self, *a, **k):
        self.args = a


class _Inert:
    def __getattr__(self, name):
        return _Inert()

    def __call__(self, *a, **k):
        return _Inert()


font = _Inert()
draw = _Inert()
image = _Inert()
display = _Inert()
Color = _Inert()

"""Inert stand-in for ujson (absent here; woker/sl.py:12 imports it, only write_game_data_to_file uses it)."""
from json import dump, dumps, load, loads  # noqa: F401

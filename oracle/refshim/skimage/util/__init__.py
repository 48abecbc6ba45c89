from ._map_array import map_array, ArrayMap  # noqa: F401

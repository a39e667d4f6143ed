"""Static map holder.  Mirrors PolygonObstacle (*/sub_systems/obstacle.py:92-124): a list of polygons
given as (east, north) vertex lists plus the map bounding box.  The point-in-polygon and
ring-distance queries (obstacle.py:126-141, Shapely in the reference) run in the CUDA kernel with
the vertices staged in shared memory."""
from __future__ import annotations


class PolygonObstacle:
    def __init__(self, list_of_vertices_list):
        self.vertices = [[(float(e), float(n)) for (e, n) in poly] for poly in list_of_vertices_list]
        self.num_obstacles = len(self.vertices)
        self.map_boundaries(list_of_vertices_list)

    def map_boundaries(self, list_of_vertices_list):
        all_points = [point for island in list_of_vertices_list for point in island]
        east_values = [p[0] for p in all_points]
        north_values = [p[1] for p in all_points]
        self.min_east = min(east_values)
        self.max_east = max(east_values)
        self.min_north = min(north_values)
        self.max_north = max(north_values)

#!/usr/bin/env python
"""Marching-cubes case tables (classic Lorensen/Bourke 256-case tables, the ones PyMCubes ships in
``mcubes/src/marchingcubes.cpp``; PyMCubes is the third-party module the reference calls at
``src/NPHM/utils/reconstruction.py:30`` and is NOT vendored in the reference checkout).

The triangle table below was written down from the published table and is validated here, row by
row, without any external source:

  * every row uses exactly the cube edges whose end corners lie on opposite sides (== edge table),
  * the triangles of a row form an oriented 2-manifold patch whose boundary lies on cube faces and,
    on every face, joins the crossed edges of that face in pairs,
  * orientation is consistent across all 256 rows (normal points from 'set' corners to 'unset'),
  * face ambiguity is resolved identically from both sides of a face (=> watertight meshes).

Running this file regenerates ``nphm_b200/csrc/mc_tables.h`` and ``oracle/mc_tables_oracle.h``.
Conventions: corner m at (x+(m in 1,2,5,6), y+(m in 2,3,6,7), z+(m>=4)); edge e joins EDGE_CORNERS[e].
"""
import itertools
import os
import sys

EDGE_CORNERS = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4),
                (0, 4), (1, 5), (2, 6), (3, 7)]
CORNER_POS = [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1)]

TRI = [
 [],
 [0, 8, 3],
 [0, 1, 9],
 [1, 8, 3, 9, 8, 1],
 [1, 2, 10],
 [0, 8, 3, 1, 2, 10],
 [9, 2, 10, 0, 2, 9],
 [2, 8, 3, 2, 10, 8, 10, 9, 8],
 [3, 11, 2],
 [0, 11, 2, 8, 11, 0],
 [1, 9, 0, 2, 3, 11],
 [1, 11, 2, 1, 9, 11, 9, 8, 11],
 [3, 10, 1, 11, 10, 3],
 [0, 10, 1, 0, 8, 10, 8, 11, 10],
 [3, 9, 0, 3, 11, 9, 11, 10, 9],
 [9, 8, 10, 10, 8, 11],
 [4, 7, 8],
 [4, 3, 0, 7, 3, 4],
 [0, 1, 9, 8, 4, 7],
 [4, 1, 9, 4, 7, 1, 7, 3, 1],
 [1, 2, 10, 8, 4, 7],
 [3, 4, 7, 3, 0, 4, 1, 2, 10],
 [9, 2, 10, 9, 0, 2, 8, 4, 7],
 [2, 10, 9, 2, 9, 7, 2, 7, 3, 7, 9, 4],
 [8, 4, 7, 3, 11, 2],
 [11, 4, 7, 11, 2, 4, 2, 0, 4],
 [9, 0, 1, 8, 4, 7, 2, 3, 11],
 [4, 7, 11, 9, 4, 11, 9, 11, 2, 9, 2, 1],
 [3, 10, 1, 3, 11, 10, 7, 8, 4],
 [1, 11, 10, 1, 4, 11, 1, 0, 4, 7, 11, 4],
 [4, 7, 8, 9, 0, 11, 9, 11, 10, 11, 0, 3],
 [4, 7, 11, 4, 11, 9, 9, 11, 10],
 [9, 5, 4],
 [9, 5, 4, 0, 8, 3],
 [0, 5, 4, 1, 5, 0],
 [8, 5, 4, 8, 3, 5, 3, 1, 5],
 [1, 2, 10, 9, 5, 4],
 [3, 0, 8, 1, 2, 10, 4, 9, 5],
 [5, 2, 10, 5, 4, 2, 4, 0, 2],
 [2, 10, 5, 3, 2, 5, 3, 5, 4, 3, 4, 8],
 [9, 5, 4, 2, 3, 11],
 [0, 11, 2, 0, 8, 11, 4, 9, 5],
 [0, 5, 4, 0, 1, 5, 2, 3, 11],
 [2, 1, 5, 2, 5, 8, 2, 8, 11, 4, 8, 5],
 [10, 3, 11, 10, 1, 3, 9, 5, 4],
 [4, 9, 5, 0, 8, 1, 8, 10, 1, 8, 11, 10],
 [5, 4, 0, 5, 0, 11, 5, 11, 10, 11, 0, 3],
 [5, 4, 8, 5, 8, 10, 10, 8, 11],
 [9, 7, 8, 5, 7, 9],
 [9, 3, 0, 9, 5, 3, 5, 7, 3],
 [0, 7, 8, 0, 1, 7, 1, 5, 7],
 [1, 5, 3, 3, 5, 7],
 [9, 7, 8, 9, 5, 7, 10, 1, 2],
 [10, 1, 2, 9, 5, 0, 5, 3, 0, 5, 7, 3],
 [8, 0, 2, 8, 2, 5, 8, 5, 7, 10, 5, 2],
 [2, 10, 5, 2, 5, 3, 3, 5, 7],
 [7, 9, 5, 7, 8, 9, 3, 11, 2],
 [9, 5, 7, 9, 7, 2, 9, 2, 0, 2, 7, 11],
 [2, 3, 11, 0, 1, 8, 1, 7, 8, 1, 5, 7],
 [11, 2, 1, 11, 1, 7, 7, 1, 5],
 [9, 5, 8, 8, 5, 7, 10, 1, 3, 10, 3, 11],
 [5, 7, 0, 5, 0, 9, 7, 11, 0, 1, 0, 10, 11, 10, 0],
 [11, 10, 0, 11, 0, 3, 10, 5, 0, 8, 0, 7, 5, 7, 0],
 [11, 10, 5, 7, 11, 5],
 [10, 6, 5],
 [0, 8, 3, 5, 10, 6],
 [9, 0, 1, 5, 10, 6],
 [1, 8, 3, 1, 9, 8, 5, 10, 6],
 [1, 6, 5, 2, 6, 1],
 [1, 6, 5, 1, 2, 6, 3, 0, 8],
 [9, 6, 5, 9, 0, 6, 0, 2, 6],
 [5, 9, 8, 5, 8, 2, 5, 2, 6, 3, 2, 8],
 [2, 3, 11, 10, 6, 5],
 [11, 0, 8, 11, 2, 0, 10, 6, 5],
 [0, 1, 9, 2, 3, 11, 5, 10, 6],
 [5, 10, 6, 1, 9, 2, 9, 11, 2, 9, 8, 11],
 [6, 3, 11, 6, 5, 3, 5, 1, 3],
 [0, 8, 11, 0, 11, 5, 0, 5, 1, 5, 11, 6],
 [3, 11, 6, 0, 3, 6, 0, 6, 5, 0, 5, 9],
 [6, 5, 9, 6, 9, 11, 11, 9, 8],
 [5, 10, 6, 4, 7, 8],
 [4, 3, 0, 4, 7, 3, 6, 5, 10],
 [1, 9, 0, 5, 10, 6, 8, 4, 7],
 [10, 6, 5, 1, 9, 7, 1, 7, 3, 7, 9, 4],
 [6, 1, 2, 6, 5, 1, 4, 7, 8],
 [1, 2, 5, 5, 2, 6, 3, 0, 4, 3, 4, 7],
 [8, 4, 7, 9, 0, 5, 0, 6, 5, 0, 2, 6],
 [7, 3, 9, 7, 9, 4, 3, 2, 9, 5, 9, 6, 2, 6, 9],
 [3, 11, 2, 7, 8, 4, 10, 6, 5],
 [5, 10, 6, 4, 7, 2, 4, 2, 0, 2, 7, 11],
 [0, 1, 9, 4, 7, 8, 2, 3, 11, 5, 10, 6],
 [9, 2, 1, 9, 11, 2, 9, 4, 11, 7, 11, 4, 5, 10, 6],
 [8, 4, 7, 3, 11, 5, 3, 5, 1, 5, 11, 6],
 [5, 1, 11, 5, 11, 6, 1, 0, 11, 7, 11, 4, 0, 4, 11],
 [0, 5, 9, 0, 6, 5, 0, 3, 6, 11, 6, 3, 8, 4, 7],
 [6, 5, 9, 6, 9, 11, 4, 7, 9, 7, 11, 9],
 [10, 4, 9, 6, 4, 10],
 [4, 10, 6, 4, 9, 10, 0, 8, 3],
 [10, 0, 1, 10, 6, 0, 6, 4, 0],
 [8, 3, 1, 8, 1, 6, 8, 6, 4, 6, 1, 10],
 [1, 4, 9, 1, 2, 4, 2, 6, 4],
 [3, 0, 8, 1, 2, 9, 2, 4, 9, 2, 6, 4],
 [0, 2, 4, 4, 2, 6],
 [8, 3, 2, 8, 2, 4, 4, 2, 6],
 [10, 4, 9, 10, 6, 4, 11, 2, 3],
 [0, 8, 2, 2, 8, 11, 4, 9, 10, 4, 10, 6],
 [3, 11, 2, 0, 1, 6, 0, 6, 4, 6, 1, 10],
 [6, 4, 1, 6, 1, 10, 4, 8, 1, 2, 1, 11, 8, 11, 1],
 [9, 6, 4, 9, 3, 6, 9, 1, 3, 11, 6, 3],
 [8, 11, 1, 8, 1, 0, 11, 6, 1, 9, 1, 4, 6, 4, 1],
 [3, 11, 6, 3, 6, 0, 0, 6, 4],
 [6, 4, 8, 11, 6, 8],
 [7, 10, 6, 7, 8, 10, 8, 9, 10],
 [0, 7, 3, 0, 10, 7, 0, 9, 10, 6, 7, 10],
 [10, 6, 7, 1, 10, 7, 1, 7, 8, 1, 8, 0],
 [10, 6, 7, 10, 7, 1, 1, 7, 3],
 [1, 2, 6, 1, 6, 8, 1, 8, 9, 8, 6, 7],
 [2, 6, 9, 2, 9, 1, 6, 7, 9, 0, 9, 3, 7, 3, 9],
 [7, 8, 0, 7, 0, 6, 6, 0, 2],
 [7, 3, 2, 6, 7, 2],
 [2, 3, 11, 10, 6, 8, 10, 8, 9, 8, 6, 7],
 [2, 0, 7, 2, 7, 11, 0, 9, 7, 6, 7, 10, 9, 10, 7],
 [1, 8, 0, 1, 7, 8, 1, 10, 7, 6, 7, 10, 2, 3, 11],
 [11, 2, 1, 11, 1, 7, 10, 6, 1, 6, 7, 1],
 [8, 9, 6, 8, 6, 7, 9, 1, 6, 11, 6, 3, 1, 3, 6],
 [0, 9, 1, 11, 6, 7],
 [7, 8, 0, 7, 0, 6, 3, 11, 0, 11, 6, 0],
 [7, 11, 6],
 [7, 6, 11],
 [3, 0, 8, 11, 7, 6],
 [0, 1, 9, 11, 7, 6],
 [8, 1, 9, 8, 3, 1, 11, 7, 6],
 [10, 1, 2, 6, 11, 7],
 [1, 2, 10, 3, 0, 8, 6, 11, 7],
 [2, 9, 0, 2, 10, 9, 6, 11, 7],
 [6, 11, 7, 2, 10, 3, 10, 8, 3, 10, 9, 8],
 [7, 2, 3, 6, 2, 7],
 [7, 0, 8, 7, 6, 0, 6, 2, 0],
 [2, 7, 6, 2, 3, 7, 0, 1, 9],
 [1, 6, 2, 1, 8, 6, 1, 9, 8, 8, 7, 6],
 [10, 7, 6, 10, 1, 7, 1, 3, 7],
 [10, 7, 6, 1, 7, 10, 1, 8, 7, 1, 0, 8],
 [0, 3, 7, 0, 7, 10, 0, 10, 9, 6, 10, 7],
 [7, 6, 10, 7, 10, 8, 8, 10, 9],
 [6, 8, 4, 11, 8, 6],
 [3, 6, 11, 3, 0, 6, 0, 4, 6],
 [8, 6, 11, 8, 4, 6, 9, 0, 1],
 [9, 4, 6, 9, 6, 3, 9, 3, 1, 11, 3, 6],
 [6, 8, 4, 6, 11, 8, 2, 10, 1],
 [1, 2, 10, 3, 0, 11, 0, 6, 11, 0, 4, 6],
 [4, 11, 8, 4, 6, 11, 0, 2, 9, 2, 10, 9],
 [10, 9, 3, 10, 3, 2, 9, 4, 3, 11, 3, 6, 4, 6, 3],
 [8, 2, 3, 8, 4, 2, 4, 6, 2],
 [0, 4, 2, 4, 6, 2],
 [1, 9, 0, 2, 3, 4, 2, 4, 6, 4, 3, 8],
 [1, 9, 4, 1, 4, 2, 2, 4, 6],
 [8, 1, 3, 8, 6, 1, 8, 4, 6, 6, 10, 1],
 [10, 1, 0, 10, 0, 6, 6, 0, 4],
 [4, 6, 3, 4, 3, 8, 6, 10, 3, 0, 3, 9, 10, 9, 3],
 [10, 9, 4, 6, 10, 4],
 [4, 9, 5, 7, 6, 11],
 [0, 8, 3, 4, 9, 5, 11, 7, 6],
 [5, 0, 1, 5, 4, 0, 7, 6, 11],
 [11, 7, 6, 8, 3, 4, 3, 5, 4, 3, 1, 5],
 [9, 5, 4, 10, 1, 2, 7, 6, 11],
 [6, 11, 7, 1, 2, 10, 0, 8, 3, 4, 9, 5],
 [7, 6, 11, 5, 4, 10, 4, 2, 10, 4, 0, 2],
 [3, 4, 8, 3, 5, 4, 3, 2, 5, 10, 5, 2, 11, 7, 6],
 [7, 2, 3, 7, 6, 2, 5, 4, 9],
 [9, 5, 4, 0, 8, 6, 0, 6, 2, 6, 8, 7],
 [3, 6, 2, 3, 7, 6, 1, 5, 0, 5, 4, 0],
 [6, 2, 8, 6, 8, 7, 2, 1, 8, 4, 8, 5, 1, 5, 8],
 [9, 5, 4, 10, 1, 6, 1, 7, 6, 1, 3, 7],
 [1, 6, 10, 1, 7, 6, 1, 0, 7, 8, 7, 0, 9, 5, 4],
 [4, 0, 10, 4, 10, 5, 0, 3, 10, 6, 10, 7, 3, 7, 10],
 [7, 6, 10, 7, 10, 8, 5, 4, 10, 4, 8, 10],
 [6, 9, 5, 6, 11, 9, 11, 8, 9],
 [3, 6, 11, 0, 6, 3, 0, 5, 6, 0, 9, 5],
 [0, 11, 8, 0, 5, 11, 0, 1, 5, 5, 6, 11],
 [6, 11, 3, 6, 3, 5, 5, 3, 1],
 [1, 2, 10, 9, 5, 11, 9, 11, 8, 11, 5, 6],
 [0, 11, 3, 0, 6, 11, 0, 9, 6, 5, 6, 9, 1, 2, 10],
 [11, 8, 5, 11, 5, 6, 8, 0, 5, 10, 5, 2, 0, 2, 5],
 [6, 11, 3, 6, 3, 5, 2, 10, 3, 10, 5, 3],
 [5, 8, 9, 5, 2, 8, 5, 6, 2, 3, 8, 2],
 [9, 5, 6, 9, 6, 0, 0, 6, 2],
 [1, 5, 8, 1, 8, 0, 5, 6, 8, 3, 8, 2, 6, 2, 8],
 [1, 5, 6, 2, 1, 6],
 [1, 3, 6, 1, 6, 10, 3, 8, 6, 5, 6, 9, 8, 9, 6],
 [10, 1, 0, 10, 0, 6, 9, 5, 0, 5, 6, 0],
 [0, 3, 8, 5, 6, 10],
 [10, 5, 6],
 [11, 5, 10, 7, 5, 11],
 [11, 5, 10, 11, 7, 5, 8, 3, 0],
 [5, 11, 7, 5, 10, 11, 1, 9, 0],
 [10, 7, 5, 10, 11, 7, 9, 8, 1, 8, 3, 1],
 [11, 1, 2, 11, 7, 1, 7, 5, 1],
 [0, 8, 3, 1, 2, 7, 1, 7, 5, 7, 2, 11],
 [9, 7, 5, 9, 2, 7, 9, 0, 2, 2, 11, 7],
 [7, 5, 2, 7, 2, 11, 5, 9, 2, 3, 2, 8, 9, 8, 2],
 [2, 5, 10, 2, 3, 5, 3, 7, 5],
 [8, 2, 0, 8, 5, 2, 8, 7, 5, 10, 2, 5],
 [9, 0, 1, 5, 10, 3, 5, 3, 7, 3, 10, 2],
 [9, 8, 2, 9, 2, 1, 8, 7, 2, 10, 2, 5, 7, 5, 2],
 [1, 3, 5, 3, 7, 5],
 [0, 8, 7, 0, 7, 1, 1, 7, 5],
 [9, 0, 3, 9, 3, 5, 5, 3, 7],
 [9, 8, 7, 5, 9, 7],
 [5, 8, 4, 5, 10, 8, 10, 11, 8],
 [5, 0, 4, 5, 11, 0, 5, 10, 11, 11, 3, 0],
 [0, 1, 9, 8, 4, 10, 8, 10, 11, 10, 4, 5],
 [10, 11, 4, 10, 4, 5, 11, 3, 4, 9, 4, 1, 3, 1, 4],
 [2, 5, 1, 2, 8, 5, 2, 11, 8, 4, 5, 8],
 [0, 4, 11, 0, 11, 3, 4, 5, 11, 2, 11, 1, 5, 1, 11],
 [0, 2, 5, 0, 5, 9, 2, 11, 5, 4, 5, 8, 11, 8, 5],
 [9, 4, 5, 2, 11, 3],
 [2, 5, 10, 3, 5, 2, 3, 4, 5, 3, 8, 4],
 [5, 10, 2, 5, 2, 4, 4, 2, 0],
 [3, 10, 2, 3, 5, 10, 3, 8, 5, 4, 5, 8, 0, 1, 9],
 [5, 10, 2, 5, 2, 4, 1, 9, 2, 9, 4, 2],
 [8, 4, 5, 8, 5, 3, 3, 5, 1],
 [0, 4, 5, 1, 0, 5],
 [8, 4, 5, 8, 5, 3, 9, 0, 5, 0, 3, 5],
 [9, 4, 5],
 [4, 11, 7, 4, 9, 11, 9, 10, 11],
 [0, 8, 3, 4, 9, 7, 9, 11, 7, 9, 10, 11],
 [1, 10, 11, 1, 11, 4, 1, 4, 0, 7, 4, 11],
 [3, 1, 4, 3, 4, 8, 1, 10, 4, 7, 4, 11, 10, 11, 4],
 [4, 11, 7, 9, 11, 4, 9, 2, 11, 9, 1, 2],
 [9, 7, 4, 9, 11, 7, 9, 1, 11, 2, 11, 1, 0, 8, 3],
 [11, 7, 4, 11, 4, 2, 2, 4, 0],
 [11, 7, 4, 11, 4, 2, 8, 3, 4, 3, 2, 4],
 [2, 9, 10, 2, 7, 9, 2, 3, 7, 7, 4, 9],
 [9, 10, 7, 9, 7, 4, 10, 2, 7, 8, 7, 0, 2, 0, 7],
 [3, 7, 10, 3, 10, 2, 7, 4, 10, 1, 10, 0, 4, 0, 10],
 [1, 10, 2, 8, 7, 4],
 [4, 9, 1, 4, 1, 7, 7, 1, 3],
 [4, 9, 1, 4, 1, 7, 0, 8, 1, 8, 7, 1],
 [4, 0, 3, 7, 4, 3],
 [4, 8, 7],
 [9, 10, 8, 10, 11, 8],
 [3, 0, 9, 3, 9, 11, 11, 9, 10],
 [0, 1, 10, 0, 10, 8, 8, 10, 11],
 [3, 1, 10, 11, 3, 10],
 [1, 2, 11, 1, 11, 9, 9, 11, 8],
 [3, 0, 9, 3, 9, 11, 1, 2, 9, 2, 11, 9],
 [0, 2, 11, 8, 0, 11],
 [3, 2, 11],
 [2, 3, 8, 2, 8, 10, 10, 8, 9],
 [9, 10, 2, 0, 9, 2],
 [2, 3, 8, 2, 8, 10, 0, 1, 8, 1, 10, 8],
 [1, 10, 2],
 [1, 3, 8, 9, 1, 8],
 [0, 9, 1],
 [0, 3, 8],
 [],
]


def edge_mask(case):
    m = 0
    for e, (a, b) in enumerate(EDGE_CORNERS):
        if ((case >> a) & 1) != ((case >> b) & 1):
            m |= 1 << e
    return m


def edge_faces(e):
    """cube faces (axis, side) that contain edge e"""
    a, b = EDGE_CORNERS[e]
    pa, pb = CORNER_POS[a], CORNER_POS[b]
    return {(ax, pa[ax]) for ax in range(3) if pa[ax] == pb[ax]}


def edge_mid(e):
    a, b = EDGE_CORNERS[e]
    return tuple((CORNER_POS[a][i] + CORNER_POS[b][i]) / 2.0 for i in range(3))


def validate():
    errors = []
    assert len(TRI) == 256, len(TRI)
    face_pairings = {}
    for case, row in enumerate(TRI):
        tag = 'case %d' % case
        if len(row) % 3 or len(row) > 15 or any(not (0 <= e < 12) for e in row):
            errors.append(tag + ': malformed row'); continue
        used = 0
        for e in row:
            used |= 1 << e
        if used != edge_mask(case):
            errors.append(tag + ': uses edges %03x but crossed edges are %03x' % (used, edge_mask(case)))
            continue
        tris = [tuple(row[i:i + 3]) for i in range(0, len(row), 3)]
        if any(len(set(t)) != 3 for t in tris):
            errors.append(tag + ': degenerate triangle'); continue
        # directed half edges
        half = {}
        for t in tris:
            for i in range(3):
                h = (t[i], t[(i + 1) % 3])
                if h in half:
                    errors.append(tag + ': half edge %s twice (inconsistent orientation)' % (h,))
                half[h] = t
        boundary = [h for h in half if (h[1], h[0]) not in half]
        # boundary half edges must lie on a cube face
        per_face = {}
        ok = True
        for (a, b) in boundary:
            common = edge_faces(a) & edge_faces(b)
            if len(common) != 1:
                errors.append(tag + ': boundary segment %d-%d not on a single cube face' % (a, b)); ok = False
                continue
            per_face.setdefault(next(iter(common)), []).append((a, b))
        if not ok:
            continue
        # on each face: crossed edges of that face are joined in pairs, each exactly once
        for face in itertools.product(range(3), (0, 1)):
            f_edges = [e for e in range(12) if face in edge_faces(e) and (edge_mask(case) >> e) & 1]
            segs = per_face.get(face, [])
            touched = sorted(x for s in segs for x in s)
            if touched != sorted(f_edges):
                errors.append(tag + ': face %s joins %s but crossed edges are %s' % (face, touched, f_edges))
            corners_on_face = tuple(sorted((c, (case >> c) & 1) for c in range(8) if CORNER_POS[c][face[0]] == face[1]))
            key = (face[0], corners_on_face_key(case, face))
            pairing = frozenset(frozenset(s) for s in segs)
            face_pairings.setdefault(key, set()).add((pairing_key(pairing, face), case))
        # orientation, tested on the boundary: walking a boundary half edge a->b on a cube face seen
        # from outside the cube, the 'set' corner of cube edge a must always lie on the same side.
        for (a, b) in boundary:
            (ax, side) = next(iter(edge_faces(a) & edge_faces(b)))
            nf = [0.0, 0.0, 0.0]
            nf[ax] = 1.0 if side == 1 else -1.0
            pa, pb = edge_mid(a), edge_mid(b)
            seg = [pb[i] - pa[i] for i in range(3)]
            left = (nf[1] * seg[2] - nf[2] * seg[1], nf[2] * seg[0] - nf[0] * seg[2], nf[0] * seg[1] - nf[1] * seg[0])
            c0, c1 = EDGE_CORNERS[a]
            inside = c0 if (case >> c0) & 1 else c1
            d = sum(left[i] * (CORNER_POS[inside][i] - pa[i]) for i in range(3))
            ORIENT.append((case, (a, b), d))
    signs = {(d > 0) for (_, _, d) in ORIENT if abs(d) > 1e-12}
    if len(signs) != 1:
        bad = [c for (c, t, d) in ORIENT if d > 0]
        good = [c for (c, t, d) in ORIENT if d < 0]
        minority = bad if len(bad) < len(good) else good
        errors.append('orientation inconsistent in cases %s' % sorted(set(minority)))
    # face ambiguity must be resolved the same way seen from both cells sharing the face: the pairing
    # of crossed edges on a face may depend only on the 4 corner states of that face.
    for key, vals in face_pairings.items():
        pk = {v[0] for v in vals}
        if len(pk) != 1:
            errors.append('face pattern %s paired differently in cases %s' % (key, sorted(v[1] for v in vals)))
    return errors


ORIENT = []


def corners_on_face_key(case, face):
    """4 corner bits of the face in a cell-independent order (by the two in-face coordinates)."""
    ax, side = face
    others = [i for i in range(3) if i != ax]
    bits = []
    for c in range(8):
        if CORNER_POS[c][ax] == side:
            bits.append((CORNER_POS[c][others[0]], CORNER_POS[c][others[1]], (case >> c) & 1))
    return tuple(sorted(bits))


def pairing_key(pairing, face):
    """pairing of crossed edges expressed in in-face coordinates (so that the face x=1 of one cell
    compares equal with the face x=0 of its neighbour)."""
    ax, side = face
    others = [i for i in range(3) if i != ax]
    out = []
    for seg in pairing:
        pts = []
        for e in seg:
            m = edge_mid(e)
            pts.append((m[others[0]], m[others[1]]))
        out.append(tuple(sorted(pts)))
    return tuple(sorted(out))


def emit(path, guard, device_qualifier):
    n_tri = [len(r) // 3 for r in TRI]
    with open(path, 'w') as f:
        f.write('// GENERATED by tools/gen_mc_tables.py - do not edit.\n')
        f.write('// Classic 256-case marching-cubes tables (edge mask, triangle list, triangle count).\n')
        f.write('#ifndef %s\n#define %s\n\n' % (guard, guard))
        f.write('%sconst unsigned short MC_EDGE_TABLE[256] = {\n' % device_qualifier)
        for i in range(0, 256, 8):
            f.write('  ' + ', '.join('0x%03x' % edge_mask(c) for c in range(i, i + 8)) + ',\n')
        f.write('};\n\n')
        f.write('%sconst signed char MC_TRI_TABLE[256][16] = {\n' % device_qualifier)
        for r in TRI:
            row = list(r) + [-1] * (16 - len(r))
            f.write('  {' + ', '.join('%2d' % v for v in row) + '},\n')
        f.write('};\n\n')
        f.write('%sconst unsigned char MC_NUM_TRIS[256] = {\n' % device_qualifier)
        for i in range(0, 256, 16):
            f.write('  ' + ', '.join('%d' % n for n in n_tri[i:i + 16]) + ',\n')
        f.write('};\n\n#endif\n')


if __name__ == '__main__':
    errs = validate()
    for e in errs:
        print('ERROR', e)
    if errs:
        sys.exit(1)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    emit(os.path.join(root, 'nphm_b200', 'csrc', 'mc_tables.h'), 'NPHM_B200_MC_TABLES_H', 'static ')
    emit(os.path.join(root, 'oracle', 'mc_tables_oracle.h'), 'NPHM_ORACLE_MC_TABLES_H', 'static ')
    print('tables valid: 256 cases, %d triangles total' % sum(len(r) // 3 for r in TRI))

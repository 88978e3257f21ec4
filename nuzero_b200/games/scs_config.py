"""SCS scenario loader: NuZero's game-config YAML schema (Games/SCS/Game_configs/*.yml, parsed by
SCS_Game.load_game_from_config, Games/SCS/SCS_Game.py:1570-1777) -> the flat tables the device
kernels read.

Randomised maps / victory points draw from numpy's legacy global generator in the same order as the
reference (Map section, then Victory_points, in file order, after `np.random.seed(seed)` when the
seed is truthy), so a (file, seed) pair names the same scenario on both sides.
"""
import math

import numpy as np
import yaml

MAGIC = 0x53435331


class ScsScenario:
    """One YAML file + one or more seeds (= maps).  All maps share board size, units and schedule."""

    def __init__(self, path, seeds=(None,)):
        with open(path) as f:
            self.raw = yaml.safe_load(f)
        self.path = path
        self.seeds = list(seeds)
        self.maps = []  # per seed: (terrain ids [RC], vp tiles p1, vp tiles p2)
        first = True
        for seed in self.seeds:
            self._parse(seed, first)
            first = False

    # define_board_sides (SCS_Game.py:1140-1158)
    @staticmethod
    def _sides(columns):
        if columns % 2:
            mid = columns // 2
            return mid - 1, mid + 1
        mid = columns // 2
        return max(0, mid - 2), min(columns - 1, mid + 1)

    def _parse(self, seed, first):
        if seed:
            np.random.seed(seed)
        d = self.raw
        unit_types, terrain = {}, {}
        terrain_ids = []
        tile_terrain, vps = None, None
        for key, val in d.items():
            if key == "Board_dimensions":
                rows, cols = int(val["rows"]), int(val["columns"])
                p1_last, p2_first = self._sides(cols)
            elif key == "Turns":
                turns = int(val)
            elif key == "Stacking_limit":
                S = int(val)
            elif key == "Units":
                for props in val.values():
                    unit_types[props["id"]] = (props["attack"], props["defense"], props["movement"])
            elif key == "Reinforcements":
                reinf = val
            elif key == "Terrain":
                for props in val.values():
                    terrain[props["id"]] = (props["attack_modifier"], props["defense_modifier"], props["cost"])
                terrain_ids = list(terrain)
            elif key == "Map":
                if val["creation_method"] == "Randomized":
                    dist = val.get("distribution") or [1 / len(terrain_ids)] * len(terrain_ids)
                    tile_terrain = [int(np.random.choice(len(terrain_ids), p=dist)) for _ in range(rows * cols)]
                elif val["creation_method"] == "Detailed":
                    grid = val["map_configuration"]
                    if np.shape(grid) != (rows, cols):
                        raise Exception("Wrong shape for map configuration, when loading game config.")
                    tile_terrain = [terrain_ids.index(grid[i][j]) for i in range(rows) for j in range(cols)]
                else:
                    raise Exception("Unrecognized map creation method, when loading game config.")
            elif key == "Victory_points":
                if val["creation_method"] == "Randomized":
                    vps = [[], []]
                    for p, n, lo, hi in ((0, val["number_vp"]["p1"], 0, p1_last + 1),
                                         (1, val["number_vp"]["p2"], p2_first, cols)):
                        if n > rows * (hi - lo):
                            raise Exception("Game config has too many victory points for p%d." % (p + 1))
                        for _ in range(n):
                            while True:
                                pt = (int(np.random.choice(range(rows))), int(np.random.choice(range(lo, hi))))
                                if pt not in vps[p]:
                                    break
                            vps[p].append(pt)
                elif val["creation_method"] == "Detailed":
                    vps = [[tuple(q) for q in val["vp_locations"]["p1"]], [tuple(q) for q in val["vp_locations"]["p2"]]]
                    for lst in vps:
                        if any(len(q) != 2 for q in lst) or len(set(lst)) != len(lst):
                            raise Exception("Bad victory point list. (game config)")
                else:
                    raise Exception("Unrecognized victory points creation method. (game config)")
        self.maps.append((tile_terrain, [r * cols + c for r, c in vps[0]], [r * cols + c for r, c in vps[1]]))
        if not first:
            return
        self.rows, self.cols, self.turns, self.S = rows, cols, turns, S
        self.terrain_types = [terrain[i] for i in terrain_ids]
        # reinforcement schedule (SCS_Game.py:1606-1656)
        method = reinf["arrival"]["method"]
        arr_sets, arr_index = [], {}

        def arrset(tiles):
            key = tuple(tiles)
            if key not in arr_index:
                arr_index[key] = len(arr_sets)
                arr_sets.append(key)
            return arr_index[key]

        default = [[i * cols + j for i in range(rows) for j in range(cols) if j <= p1_last],
                   [i * cols + j for i in range(rows) for j in range(cols) if j >= p2_first]]
        counters = [0, 0]
        units = [[], []]
        for pname, per_turn in reinf["schedule"].items():
            if len(per_turn) != turns + 1:
                raise Exception("Reinforcement schedule should have 'turns + 1' entries.")
            p = int(pname[-1]) - 1
            for turn, ids in enumerate(per_turn):
                for uid in ids:
                    if method == "Default":
                        tiles = default[p]
                    else:
                        tiles = [q[0] * cols + q[1] for q in reinf["arrival"]["locations"]["p%d" % (p + 1)][counters[p]]]
                        counters[p] += 1
                    a, dfn, mv = unit_types[uid]
                    units[p].append((p, turn, a, dfn, mv, arrset(tiles)))
        self.units = units[0] + units[1]
        self.count = (len(units[0]), len(units[1]))
        self.arr_sets = arr_sets
        self.planes = 3 + 9 * S
        self.action_shape = (self.planes, rows, cols)
        self.A = self.planes * rows * cols
        self.C = 48 + 19 * S
        self.state_shape = (self.C, rows, cols)

    def to_desc(self):
        RC = self.rows * self.cols
        nv0, nv1 = len(self.maps[0][1]), len(self.maps[0][2])
        for m in self.maps:
            if (len(m[1]), len(m[2])) != (nv0, nv1):
                raise Exception("all maps of a scenario need the same number of victory points")
        for u in self.units:
            if any(int(x) != x for x in u):
                raise Exception("unit stats must be integers")
        out = [MAGIC, self.rows, self.cols, self.turns, self.S, len(self.terrain_types), len(self.units),
               len(self.maps), nv0, nv1, self.count[0], self.count[1], len(self.arr_sets), 0, 0, 0]
        types = np.array(self.terrain_types, dtype=np.float64).reshape(-1)
        out += types.view(np.int32).tolist()
        for u in self.units:
            out += [int(x) for x in u]
        words = (RC + 31) // 32
        for tiles in self.arr_sets:
            bits = np.zeros(words * 32, dtype=np.uint8)
            bits[list(tiles)] = 1
            out += np.packbits(bits.reshape(words, 32), axis=1, bitorder="little").view("<u4").reshape(-1).astype(np.int64).tolist()
        for m in self.maps:
            out += [int(x) for x in m[0]]
        for m in self.maps:
            out += [int(x) for x in m[1]] + [int(x) for x in m[2]]
        arr = np.array(out, dtype=np.int64)
        arr = np.where(arr >= 2 ** 31, arr - 2 ** 32, arr)
        return arr.astype(np.int32)

    def spec(self, max_moves=600):
        from .. import _ffi
        from ..engine import GameSpec

        RC = self.rows * self.cols
        max_children = min(self.A, RC + 8 * max(self.count) + 2)
        return GameSpec(_ffi.GAME_SCS, self.to_desc(), max_children=max_children, max_moves=max_moves,
                        name="scs:%s" % self.path)

"""TEST INFRASTRUCTURE — the SCS hex wargame restated from Games/SCS/SCS_Game.py (+ Unit.py, Tile.py,
Terrain.py) on flat tables: units are rows of small integer arrays, a tile is an ordered list of
unit ids, the three per-player status lists of the reference become one status field per unit
(SURVEY.md Appendix A explains why that is sufficient on the search path).

Everything is indexed by `tile = row * columns + col`.  Players are 0 and 1 (ref :93).
"""
import math

import numpy as np
import yaml

QUEUED, AVAILABLE, MOVED, ATTACKED, DEAD = 3, 0, 1, 2, 4


# --------------------------------------------------------------------------------------------
# scenario = everything load_game_from_config (ref :1570-1777) derives from the YAML + seed
# --------------------------------------------------------------------------------------------
class Scenario:
    pass


def board_sides(columns):
    """define_board_sides (ref :1140-1158) -> (p1_last_index, p2_first_index)."""
    if columns % 2 != 0:
        mid = math.floor(columns / 2)
        return mid - 1, mid + 1
    mid = int(columns / 2)
    return max(0, (mid - 1) - 1), min(columns - 1, mid + 1)


def load_scenario(path, seed=None):
    with open(path) as f:
        data = yaml.safe_load(f)
    if seed:  # ref :1575-1576 — seed 0 / None leaves numpy's global generator untouched
        np.random.seed(seed)
    sc = Scenario()
    sc.vp = [[], []]
    units_by_id, terrain_by_id = {}, {}
    schedule = None
    for section, values in data.items():  # section order matters: it fixes the RNG draw order
        if section == "Name":
            sc.title = values
        elif section == "Board_dimensions":
            sc.rows, sc.cols = values["rows"], values["columns"]
            sc.p1_last, sc.p2_first = board_sides(sc.cols)
        elif section == "Turns":
            sc.turns = values
        elif section == "Stacking_limit":
            sc.S = values
        elif section == "Units":
            for _name, props in values.items():
                units_by_id[props["id"]] = (props["attack"], props["defense"], props["movement"])
        elif section == "Reinforcements":
            schedule, arrival = values["schedule"], values["arrival"]
            method = arrival["method"]
            if method == "Default":  # ref :1611-1619
                default_loc = [[], []]
                for i in range(sc.rows):
                    for j in range(sc.cols):
                        if j <= sc.p1_last:
                            default_loc[0].append(i * sc.cols + j)
                        elif j >= sc.p2_first:
                            default_loc[1].append(i * sc.cols + j)
            else:
                detailed = [arrival["locations"]["p1"], arrival["locations"]["p2"]]
                counters = [0, 0]
            sc.sched = [None, None]  # per player: list of (turn, (atk, def, mov), [arrival tiles])
            for pname, per_turn in schedule.items():
                if len(per_turn) != sc.turns + 1:
                    raise Exception("Reinforcement schedule should have 'turns + 1' entries.")
                p = int(pname[-1]) - 1
                rows = []
                for turn, ids in enumerate(per_turn):
                    for uid in ids:
                        if method == "Default":
                            loc = list(default_loc[p])
                        else:
                            loc = [pt[0] * sc.cols + pt[1] for pt in detailed[p][counters[p]]]
                            counters[p] += 1
                        rows.append((turn, units_by_id[uid], loc))
                sc.sched[p] = rows
        elif section == "Terrain":
            for _name, props in values.items():
                terrain_by_id[props["id"]] = (props["attack_modifier"], props["defense_modifier"], props["cost"])
        elif section == "Map":
            ids = list(terrain_by_id.keys())
            sc.terrain_types = [terrain_by_id[i] for i in ids]
            if values["creation_method"] == "Randomized":  # ref :1680-1691
                dist = values.get("distribution") or [1 / len(ids) for _ in ids]
                sc.tile_terrain = [int(np.random.choice(len(ids), p=dist)) for _ in range(sc.rows * sc.cols)]
            else:
                cfg = values["map_configuration"]
                if np.shape(cfg) != (sc.rows, sc.cols):
                    raise Exception("Wrong shape for map configuration, when loading game config.")
                sc.tile_terrain = [ids.index(cfg[i][j]) for i in range(sc.rows) for j in range(sc.cols)]
        elif section == "Victory_points":
            if values["creation_method"] == "Randomized":  # ref :1711-1744
                n1, n2 = values["number_vp"]["p1"], values["number_vp"]["p2"]
                for _ in range(n1):
                    while True:
                        pt = (int(np.random.choice(range(sc.rows))), int(np.random.choice(range(sc.p1_last + 1))))
                        if pt not in sc.vp[0]:
                            break
                    sc.vp[0].append(pt)
                for _ in range(n2):
                    while True:
                        pt = (int(np.random.choice(range(sc.rows))),
                              int(np.random.choice(range(sc.p2_first, sc.cols))))
                        if pt not in sc.vp[1]:
                            break
                    sc.vp[1].append(pt)
            else:
                sc.vp = [[tuple(p) for p in values["vp_locations"]["p1"]],
                         [tuple(p) for p in values["vp_locations"]["p2"]]]
    sc.vp_tiles = [[r * sc.cols + c for (r, c) in lst] for lst in sc.vp]
    # unit table: player 0's schedule first, then player 1's, each in (turn, position-in-turn) order
    sc.unit_player, sc.unit_turn, sc.unit_stats, sc.unit_arrival = [], [], [], []
    sc.first_unit = [0, 0]
    for p in range(2):
        sc.first_unit[p] = len(sc.unit_player)
        for (turn, stats, loc) in sc.sched[p]:
            sc.unit_player.append(p)
            sc.unit_turn.append(turn)
            sc.unit_stats.append(stats)
            sc.unit_arrival.append(loc)
    sc.n_units = len(sc.unit_player)
    sc.count = [len(sc.sched[0]), len(sc.sched[1])]
    # cum[p][t] = units of p scheduled strictly before turn t
    sc.cum = [[sum(1 for (tt, _, _) in sc.sched[p] if tt < t) for t in range(sc.turns + 2)] for p in range(2)]
    sc.planes = 3 + 9 * sc.S  # ref :150-170
    sc.A = sc.planes * sc.rows * sc.cols
    sc.C = 48 + 19 * sc.S  # ref :188-239
    return sc


# --------------------------------------------------------------------------------------------
class SCS:
    prior_is_f64 = False  # int8 mask -> float32 priors (ref :399-408, Explorer.py:168-179)

    def __init__(self, scenario, _fresh=True):
        sc = self.sc = scenario
        self.action_space_shape = (sc.planes, sc.rows, sc.cols)
        self.num_actions = sc.A
        self.state_shape = (sc.C, sc.rows, sc.cols)
        if not _fresh:
            return
        self.stage, self.turn, self.length = -2, 0, 0
        self.terminal, self.terminal_value = False, 0
        self.player, self.sub_phase = 0, 0
        self.placed = [0, 0]
        self.status = [QUEUED] * sc.n_units
        self.pos = [-1] * sc.n_units
        self.mov = [st[2] for st in sc.unit_stats]
        self.stack = [[] for _ in range(sc.rows * sc.cols)]
        self.target = -1
        self.attackers = []
        self._update_env()  # ref :287

    # -- interface used by the search -------------------------------------------------------------
    def get_current_player(self):
        return self.player

    def is_terminal(self):
        return self.terminal

    def get_terminal_value(self):
        return self.terminal_value

    def get_length(self):
        return self.length

    def get_num_actions(self):
        return self.num_actions

    def get_winner(self):  # ref :896-906
        return 2 if self.terminal_value < 0 else (1 if self.terminal_value > 0 else 0)

    def clone(self):  # shallow_clone (ref :1782-1793) copies every dynamic attribute
        c = SCS(self.sc, _fresh=False)
        c.stage, c.turn, c.length = self.stage, self.turn, self.length
        c.terminal, c.terminal_value = self.terminal, self.terminal_value
        c.player, c.sub_phase = self.player, self.sub_phase
        c.placed = list(self.placed)
        c.status, c.pos, c.mov = list(self.status), list(self.pos), list(self.mov)
        c.stack = [list(s) for s in self.stack]
        c.target, c.attackers = self.target, list(self.attackers)
        return c

    # -- geometry (ref :1048-1094, :1199-1243): even columns are shifted up ------------------------
    def neighbours(self, tile):
        """check_tiles: [n, ne, se, s, sw, nw] tile ids or -1 when off the board."""
        R, C = self.sc.rows, self.sc.cols
        r, c = divmod(tile, C)
        even = (c % 2 == 0)
        out = [-1] * 6
        if r - 1 != -1:
            out[0] = (r - 1) * C + c
        if r + 1 != R:
            out[3] = (r + 1) * C + c
        if not (c == 0 or (r == 0 and even)):
            out[5] = ((r - 1) if even else r) * C + (c - 1)
        if not (c == 0 or (r == R - 1 and not even)):
            out[4] = (r if even else (r + 1)) * C + (c - 1)
        if not (c == C - 1 or (r == 0 and even)):
            out[1] = ((r - 1) if even else r) * C + (c + 1)
        if not (c == C - 1 or (r == R - 1 and not even)):
            out[2] = (r if even else (r + 1)) * C + (c + 1)
        return out

    def _dest(self, tile, direction):
        """get_{n,ne,se,s,sw,nw}_coords without any bounds test (ref :1199-1243)."""
        C = self.sc.cols
        r, c = divmod(tile, C)
        even = (c % 2 == 0)
        dr, dc = [(-1, 0), (-1 if even else 0, 1), (0 if even else 1, 1), (1, 0), (0 if even else 1, -1),
                  (-1 if even else 0, -1)][direction]
        return (r + dr) * C + (c + dc)

    def _owner(self, tile):
        """tile.player (Tile.py:28-36): owner of the units on the tile, -1 when empty."""
        st = self.stack[tile]
        return self.sc.unit_player[st[0]] if st else -1

    def _cost(self, tile):
        return self.sc.terrain_types[self.sc.tile_terrain[tile]][2]

    def _mobility(self, u, consider_units):
        """check_mobility (ref :1096-1111)."""
        me = self.sc.unit_player[u]
        out = []
        for t in self.neighbours(self.pos[u]):
            ok = False
            if t >= 0 and self.mov[u] - self._cost(t) >= 0:
                ok = True
                if consider_units and (len(self.stack[t]) == self.sc.S or self._owner(t) == (me ^ 1)):
                    ok = False
            out.append(ok)
        return out

    def _adjacent_units(self, tile, player):
        """check_adjacent_units (ref :1113-1124)."""
        out = []
        for t in self.neighbours(tile):
            if t >= 0:
                out.extend(u for u in self.stack[t] if self.sc.unit_player[u] == player)
        return out

    def _units_with(self, player, status):
        return [u for u in range(self.sc.n_units) if self.sc.unit_player[u] == player and self.status[u] == status]

    def _queue_empty(self, p, turn):
        """player_ended_reinforcements (ref :908-911)."""
        return self.placed[p] >= self.sc.cum[p][turn + 1]

    # -- possible_actions (ref :395-484) --------------------------------------------------------------
    def legal_mask(self):
        sc = self.sc
        RC = sc.rows * sc.cols
        S = sc.S
        m = np.zeros(sc.A, dtype=np.int8)
        p = self.player
        P_MOVE, P_TARGET, P_ATT, P_CONFIRM, P_NOMOVE, P_NOFIGHT = 1, 1 + 6 * S, 2 + 6 * S, 2 + 7 * S, 3 + 7 * S, 3 + 8 * S
        if self.sub_phase == 0:
            u = sc.first_unit[p] + self.placed[p]  # get_next_reinforcement (ref :1331-1332)
            for t in sc.unit_arrival[u]:
                if not (self._owner(t) == (p ^ 1) or len(self.stack[t]) == S):
                    m[t] = 1
        elif self.sub_phase == 1:
            for u in self._units_with(p, AVAILABLE):
                t = self.pos[u]
                s = self.stack[t].index(u)
                m[(P_NOMOVE + s) * RC + t] = 1
                for d, ok in enumerate(self._mobility(u, True)):
                    if ok:
                        m[(P_MOVE + d * S + s) * RC + t] = 1
        elif self.sub_phase == 2:
            for u in self._units_with(p, MOVED):
                t = self.pos[u]
                s = self.stack[t].index(u)
                m[(P_NOFIGHT + s) * RC + t] = 1
                for e in self._adjacent_units(t, p ^ 1):
                    m[P_TARGET * RC + self.pos[e]] = 1
        elif self.sub_phase == 3:
            for u in self._adjacent_units(self.target, p):
                if u in self.attackers or self.status[u] == ATTACKED:
                    continue
                t = self.pos[u]
                m[(P_ATT + self.stack[t].index(u)) * RC + t] = 1
            if self.attackers:
                m[P_CONFIRM * RC + self.target] = 1
        else:
            raise Exception("Error in possible_actions! Exiting")
        return m

    # -- step (ref :375-391) -> play_action (:569-633) -> update_game_env (:687-831) -------------------
    def step(self, action, check=False):
        if check and not self.legal_mask()[action]:
            raise Exception("Tried to play an illegal action!")
        sc = self.sc
        RC, S = sc.rows * sc.cols, sc.S
        plane, tile = divmod(int(action), RC)
        if plane < 1:  # placement
            p = self.player
            u = sc.first_unit[p] + self.placed[p]
            self.placed[p] += 1
            self.pos[u] = tile
            self.status[u] = AVAILABLE
            self.stack[tile].append(u)
        elif plane < 1 + 6 * S:  # movement
            d, s = divmod(plane - 1, S)
            u = self.stack[tile][s]
            dest = self._dest(tile, d)
            self.mov[u] -= self._cost(dest)
            self.pos[u] = dest
            self.stack[dest].append(u)
            self.stack[tile].remove(u)
            if not any(self._mobility(u, False)):  # ref :598-599
                self._end_movement(u)
        elif plane < 2 + 6 * S:  # choose target
            self.target = tile
        elif plane < 2 + 7 * S:  # choose attacker
            self.attackers.append(self.stack[tile][plane - (2 + 6 * S)])
        elif plane < 3 + 7 * S:  # confirm attack
            self._resolve_combat()
            self.target = -1
            self.attackers = []
        elif plane < 3 + 8 * S:  # no move
            self._end_movement(self.stack[tile][plane - (3 + 7 * S)])
        elif plane < 3 + 9 * S:  # no fight
            self.status[self.stack[tile][plane - (3 + 8 * S)]] = ATTACKED
        else:
            raise Exception("Problem parsing action...Exiting")
        self.length += 1
        self._update_env()

    def _end_movement(self, u):  # ref :927-940
        self.status[u] = MOVED
        if not self._adjacent_units(self.pos[u], self.sc.unit_player[u] ^ 1):
            self.status[u] = ATTACKED

    def _destroy(self, u):  # ref :982-995
        self.stack[self.pos[u]].remove(u)
        self.status[u] = DEAD

    def _strongest(self, units, first_key, second_key):
        """get_strongest_attacker / _defender (ref :1253-1285): strict improvements only, so the
        first unit in list order wins exact ties."""
        st = self.sc.unit_stats
        best = units[0]
        for u in units:
            a, b = st[u][first_key], st[best][first_key]
            if a > b:
                best = u
            elif a == b:
                a2, b2 = st[u][second_key], st[best][second_key]
                if a2 > b2:
                    best = u
                elif a2 == b2 and st[u][2] > st[best][2]:
                    best = u
        return best

    def _resolve_combat(self):  # ref :997-1044
        sc = self.sc
        defenders = self.stack[self.target]  # live list, like the reference's tile.units
        total_def = 0
        for u in defenders:
            total_def += sc.unit_stats[u][1]
        total_def = total_def * sc.terrain_types[sc.tile_terrain[self.target]][1]
        total_att = 0
        for u in self.attackers:
            total_att += sc.unit_stats[u][0] * sc.terrain_types[sc.tile_terrain[self.pos[u]]][0]
            self.status[u] = ATTACKED
        att_loss = 1 if total_att <= total_def else 0
        def_loss = 1 if total_att >= total_def else 0
        for _ in range(att_loss):
            self._destroy(self._strongest(self.attackers, 0, 1))
        for _ in range(def_loss):
            self._destroy(self._strongest(defenders, 1, 0))

    def _update_env(self):
        stage, done = self.stage, False
        while True:
            if stage == -2:
                if self._queue_empty(0, self.turn):
                    stage += 1
                    continue
            elif stage == -1:
                if self._queue_empty(1, self.turn):
                    self.turn += 1
                    stage += 1
                    continue
            elif stage in (0, 4):
                if self._queue_empty(stage // 4, self.turn):
                    stage += 1
                    continue
            elif stage in (1, 5):
                if not self._units_with(stage // 4, AVAILABLE):
                    stage += 1
                    continue
            elif stage in (2, 6):
                p = stage // 4
                if not self._units_with(p, MOVED):
                    if p == 0:
                        stage = 4
                        continue
                    if self.turn + 1 > self.sc.turns:
                        done = True
                        break
                    self.turn += 1
                    stage = 0
                    for u in range(self.sc.n_units):  # new_turn (ref :845-855)
                        if self.status[u] == ATTACKED:
                            self.status[u] = AVAILABLE
                            self.mov[u] = self.sc.unit_stats[u][2]
                    continue
                elif self.target >= 0:
                    stage += 1
                    continue
            elif stage in (3, 7):
                if self.target < 0:
                    stage -= 1
                    continue
            break
        self.player = 0 if stage in (-2, 0, 1, 2, 3) else 1
        if done:
            self.terminal = True
            self._check_termination()
        self.sub_phase = {-2: 0, -1: 0, 0: 0, 4: 0, 1: 1, 5: 1, 2: 2, 6: 2, 3: 3, 7: 3}[stage]
        self.stage = stage

    def _check_termination(self):  # ref :857-894
        sc = self.sc
        p2_captured = sum(1 for t in sc.vp_tiles[0] if self._owner(t) == 1)
        p1_captured = sum(1 for t in sc.vp_tiles[1] if self._owner(t) == 0)
        a, b = p1_captured / len(sc.vp_tiles[1]), p2_captured / len(sc.vp_tiles[0])
        self.terminal_value = 1 if a > b else (-1 if a < b else 0)

    # -- generate_state / generate_network_input (ref :1348-1515) --------------------------------------
    def encode(self):
        sc = self.sc
        R, C, S, RC = sc.rows, sc.cols, sc.S, sc.rows * sc.cols
        out = np.zeros((sc.C, RC), dtype=np.float32)
        for t in range(RC):
            out[0:3, t] = sc.terrain_types[sc.tile_terrain[t]]
        for p in range(2):
            for t in sc.vp_tiles[p]:
                out[3 + p, t] = 1.0
        for p in range(2):  # next three queued units of each player, in schedule order
            base = 5 + 18 * p
            for k in range(3):
                idx = self.placed[p] + k
                if idx >= sc.count[p]:
                    break
                u = sc.first_unit[p] + idx
                a, d, mv = sc.unit_stats[u]
                for t in sc.unit_arrival[u]:
                    out[base + 6 * k + 0, t] = a
                    out[base + 6 * k + 1, t] = d
                    out[base + 6 * k + 2, t] = mv
                turns_left = sc.unit_turn[u] - self.turn
                out[base + 6 * k + 3: base + 6 * k + 6, :] = ((sc.turns + 1) - turns_left) / (sc.turns + 1)
        ub = 41
        for u in range(sc.n_units):
            stt = self.status[u]
            if stt in (AVAILABLE, MOVED, ATTACKED):
                t = self.pos[u]
                ch = ub + sc.unit_player[u] * 9 * S + stt * 3 * S + self.stack[t].index(u) * 3
                out[ch + 0, t] = sc.unit_stats[u][0]
                out[ch + 1, t] = sc.unit_stats[u][1]
                out[ch + 2, t] = self.mov[u]
        fb = ub + 18 * S
        if self.target >= 0:
            out[fb, self.target] = 1.0
        for u in self.attackers:
            t = self.pos[u]
            out[fb + 1 + self.stack[t].index(u), t] = 1.0
        out[fb + 1 + S + self.sub_phase, :] = 1.0
        out[fb + 5 + S, :] = self.turn / sc.turns
        out[fb + 6 + S, :] = -1.0 if self.player == 1 else 1.0
        return out.reshape(1, sc.C, R, C)

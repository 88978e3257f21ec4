"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the NuZero self-play hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs
may import this package; the product (`nuzero_b200/`) never does and fails loudly when its CUDA
extension is missing.

Parity pinning: the reference holds NO golden vectors or known-answer tests for this path
(SURVEY.md §4), so the oracle is pinned against outputs of the reference itself, generated in the
build container by `oracle/gen_golden.py` (which imports the unmodified reference through
`oracle/ref_harness.py`) and committed as fixtures under `tests/golden/`.
`tests/test_oracle_vs_golden.py` replays every fixture through this restatement.

Modules
  ttt.py        Games/Tic_Tac_Toe/tic_tac_toe.py restated on two 9-bit boards
  scs.py        Games/SCS/SCS_Game.py (+Unit/Tile/Terrain) restated on flat unit tables
  mcts.py       Search/Node.py + Search/Explorer.py restated (explicit f32/f64 arithmetic chains)
  selfplay.py   Training/Gamer.py:39-97 game loop restated, emitting comparable records
  stubnet_np.py deterministic dyadic stub network of the parity protocol
"""

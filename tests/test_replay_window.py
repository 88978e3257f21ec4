"""CPU: the dense-ring row bookkeeping of DeviceReplayBuffer against the list semantics of Training/ReplayBuffer.py:24-36
(restated in nuzero_b200.replay.ReplayBuffer), and that restatement against the reference class itself where the
reference tree is present (build container)."""
import numpy as np
import pytest

from nuzero_b200.replay import ReplayBuffer, WindowRows


class _Game:
    def __init__(self, n, tag):
        self.state_history = [(tag, i) for i in range(n)]

    def get_state_from_history(self, i):
        return self.state_history[i]

    def make_target(self, i):
        return (0, [i])


@pytest.mark.parametrize("window", [1, 3, 7])
def test_window_rows_follow_the_list_semantics(window):
    rng = np.random.default_rng(window)
    rb, wr, phys = ReplayBuffer(window, 4), WindowRows(window, 64), [None] * 64
    for g in range(40):
        n = int(rng.integers(1, 9))
        rb.save_game(_Game(n, g), 0)
        for i, r in enumerate(wr.place(n)):
            phys[r] = (g, i)
        assert [phys[r] for r in wr.logical_rows()] == [e[0] for e in rb.get_buffer()]
        assert [phys[r] for r in wr.logical_rows(2, 5)] == [e[0] for e in rb.get_slice(2, 5)]
        assert wr.n_games == rb.played_games() and wr.count == rb.len()


def test_list_restatement_matches_reference_class():
    from oracle import ref_harness

    if not ref_harness.available():
        pytest.skip("reference tree not present (GPU box)")
    ref_harness.load()
    import importlib

    mod = importlib.import_module("Training.ReplayBuffer")
    cls = mod.ReplayBuffer
    cls = getattr(cls, "__ray_actor_class__", getattr(cls, "_cls", cls))
    rng = np.random.default_rng(0)
    a, b = cls(3, 4), ReplayBuffer(3, 4)
    for g in range(12):
        game = _Game(int(rng.integers(1, 7)), g)
        a.save_game(game, g)
        b.save_game(game, g)
        assert a.get_buffer() == b.get_buffer()
        assert a.len() == b.len() and a.played_games() == b.played_games()
        assert a.get_slice(1, 4) == b.get_slice(1, 4)


def test_place_many_equals_repeated_place():
    rng = np.random.default_rng(1)
    for window in (1, 3, 7, 50):
        a, b = WindowRows(window, 997), WindowRows(window, 997)
        for _ in range(30):
            counts = rng.integers(1, 9, size=int(rng.integers(1, 6)))
            ra = np.concatenate([a.place(int(c)) for c in counts])
            assert np.array_equal(ra, b.place_many(counts))
            assert (a.start, a.count, a.n_games) == (b.start, b.count, b.n_games)


def test_late_heavy_probs_are_the_reference_ramp():
    """Training/AlphaZero.py:779-792 written out literally."""
    from nuzero_b200.replay import late_heavy_probs

    for n in (1, 2, 7, 100, 5000):
        variation = 0.5
        offset = (1 - variation) / 2
        fraction = variation / n
        probs, total = [], offset
        for _ in range(n):
            total += fraction
            probs.append(total)
        total_sum = sum(probs)
        want = np.array([p / total_sum for p in probs])
        got = late_heavy_probs(n)
        assert got.dtype == np.float64 and np.array_equal(got, want)
        assert abs(got.sum() - 1.0) < 1e-12 and (np.diff(got) > 0).all() if n > 1 else True
    assert late_heavy_probs(0).shape == (0,)


def test_get_sample_accepts_late_heavy_probs_and_the_reference_empty_list():
    """late_heavy_probs returns an ndarray; `probs != []` on it raised on numpy 2 (list API path)."""
    from nuzero_b200.replay import late_heavy_probs

    rb = ReplayBuffer(10, 4)
    for g in range(5):
        rb.save_game(_Game(3, g), g)
    np.random.seed(0)
    got = rb.get_sample(8, True, late_heavy_probs(rb.len()))
    assert len(got) == 8 and all(e in rb.get_buffer() for e in got)
    assert len(rb.get_sample(4, False, [])) == 4          # the reference's own call shape (AlphaZero.py:796)
    assert len(rb.get_sample(4, True, list(late_heavy_probs(rb.len())))) == 4

"""CPU: the torch restatement of the 7-tap hexagonal convolution against hand-computed one-hot
responses on the SCS board layout (even columns shifted up), and model shape contracts."""
import numpy as np
import torch

from nuzero_b200.nets import HexConv2d, MLP_Network, RecurrentNet, ResNet, initialize_parameters
from oracle.scs import SCS, Scenario


def _neigh(rows, cols, tile):
    sc = Scenario()
    sc.rows, sc.cols = rows, cols
    g = SCS.__new__(SCS)
    g.sc = sc
    return g.neighbours(tile)  # n, ne, se, s, sw, nw (oracle geometry, pinned against the reference)


def test_hexconv_one_hot_responses_follow_board_geometry():
    R, C = 5, 6
    conv = HexConv2d(1, 1)
    with torch.no_grad():
        conv.kernel0.copy_(torch.tensor([1.0, 2.0, 3.0]).view(1, 1, 3, 1))          # N, centre, S
        conv.kernel1.copy_(torch.tensor([[4.0, 5.0], [6.0, 7.0]]).view(1, 1, 2, 2))  # [[NW, NE], [SW, SE]]
    tap = {0: 1.0, 1: 5.0, 2: 7.0, 3: 3.0, 4: 6.0, 5: 4.0}  # direction -> weight applied to that neighbour
    for tile in range(R * C):
        # output at `tile` = sum over neighbours of weight(direction) * input(neighbour)
        for d, nt in enumerate(_neigh(R, C, tile)):
            if nt < 0:
                continue
            x = torch.zeros(1, 1, R, C)
            x.view(-1)[nt] = 1.0
            y = conv(x).view(-1)
            assert y[tile].item() == tap[d], (tile, d, nt)
        x = torch.zeros(1, 1, R, C)
        x.view(-1)[tile] = 1.0
        assert conv(x).view(-1)[tile].item() == 2.0
        # a one-hot input excites exactly itself and its on-board neighbours
        nz = set(np.flatnonzero(conv(x).detach().view(-1).numpy()).tolist())
        assert nz == {tile} | {n for n in _neigh(R, C, tile) if n >= 0}


def test_recurrent_net_shapes_and_state_dict_names():
    torch.manual_seed(0)
    net = RecurrentNet(86, 21, 32, 2, recall=True, policy_head="conv", value_head="reduce", value_activation="relu", hex=True)
    initialize_parameters(net)
    x = torch.randn(3, 86, 5, 5)
    (p, v), thought = net(x, 3)
    assert p.shape == (3, 21, 5, 5) and v.shape == (3, 1) and thought.shape == (3, 32, 5, 5)
    assert float(v.abs().max()) <= 1.0
    names = set(net.state_dict())
    assert {"projection.0.kernel0", "projection.0.kernel1", "recur_module.0.kernel0",
            "recur_module.1.before_shortcut.0.kernel0", "policy_head.layers.0.kernel0"} <= names
    (p2, _), _ = net(x, 1, interim_thought=net(x, 2)[1])
    torch.testing.assert_close(p2, p)
    ttt = RecurrentNet(2, 1, 16, 2, hex=False)
    (p, v), _ = ttt(torch.zeros(4, 2, 3, 3), 2)
    assert p.shape == (4, 1, 3, 3) and v.shape == (4, 1)
    r = ResNet(67, 12, 16, 2, hex=True)
    p, v = r(torch.zeros(2, 67, 7, 7))
    assert p.shape == (2, 12, 7, 7) and v.shape == (2, 1) and r.recurrent is False
    m = MLP_Network(9)
    p, v = m(torch.zeros(5, 2, 3, 3))
    assert p.shape == (5, 9) and v.shape == (5, 1)


def test_hexconv_seven_tap_formulation_equals_dense_formulation():
    torch.manual_seed(1)
    for (R, C) in [(5, 5), (4, 6), (3, 1), (1, 4), (7, 2), (15, 15)]:
        conv = HexConv2d(3, 5, bias=True)
        with torch.no_grad():
            conv.bias.normal_()
        x = torch.randn(2, 3, R, C, dtype=torch.float64)
        conv = conv.double()
        torch.testing.assert_close(conv(x), conv.forward_dense(x), rtol=1e-12, atol=1e-12)

"""Policy-value networks with the reference's constructor signatures and forward conventions
(Neural_Networks/Architectures/RecurrentNet.py, ResNet.py, MLP_Network.py, blocks.py), written on plain
torch so that they can be built without the `hexagdly` package, which the reference imports but does
not vendor.  The network is the one dense contraction of the path: it runs as the PyTorch forward
(bf16, captured in a CUDA graph by nuzero_b200.network.GraphedForward), not as a hand-written kernel.

HexConv2d restates the 7-tap hexagonal convolution (`hexagdly.Conv2d(kernel_size=1)`) for the board
layout of the SCS game (even columns shifted up, Games/SCS/SCS_Game.py:27-65,1199-1243):
parameters `kernel0` (out, in, 3, 1) = north / centre / south taps and `kernel1` (out, in, 2, 2) =
{upper, lower} x {left, right} side taps.  Parity of this restatement against hexagdly itself is
UNPINNED (the package is not available here); tests pin it against hand-computed one-hot responses.
"""
import torch
import torch.nn.functional as F
from torch import nn


class HexConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, bias=False):
        super().__init__()
        if kernel_size != 1 or stride != 1:
            raise ValueError("only the 7-tap (kernel_size=1, stride=1) hexagonal convolution is used by NuZero")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel0 = nn.Parameter(torch.empty(out_channels, in_channels, 3, 1))
        self.kernel1 = nn.Parameter(torch.empty(out_channels, in_channels, 2, 2))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        nn.init.xavier_uniform_(self.kernel0)
        nn.init.xavier_uniform_(self.kernel1)

    def dense_kernels(self):
        """The two 3x3 cross-correlation kernels (for outputs on even / odd columns)."""
        k0, k1 = self.kernel0, self.kernel1
        even = k0.new_zeros(self.out_channels, self.in_channels, 3, 3)
        odd = k0.new_zeros(self.out_channels, self.in_channels, 3, 3)
        even[:, :, :, 1] = k0[:, :, :, 0]
        odd[:, :, :, 1] = k0[:, :, :, 0]
        # even column (r, c): NW (r-1, c-1), SW (r, c-1), NE (r-1, c+1), SE (r, c+1)
        even[:, :, 0, 0], even[:, :, 1, 0] = k1[:, :, 0, 0], k1[:, :, 1, 0]
        even[:, :, 0, 2], even[:, :, 1, 2] = k1[:, :, 0, 1], k1[:, :, 1, 1]
        # odd column (r, c): NW (r, c-1), SW (r+1, c-1), NE (r, c+1), SE (r+1, c+1)
        odd[:, :, 1, 0], odd[:, :, 2, 0] = k1[:, :, 0, 0], k1[:, :, 1, 0]
        odd[:, :, 1, 2], odd[:, :, 2, 2] = k1[:, :, 0, 1], k1[:, :, 1, 1]
        return even, odd

    def forward_dense(self, x):
        """Reference formulation: two zero-padded 3x3 kernels, selected by column parity (18 taps)."""
        even, odd = self.dense_kernels()
        w = torch.cat([even, odd], 0)
        y = F.conv2d(x, w, None, padding=1)
        ye, yo = y[:, : self.out_channels], y[:, self.out_channels:]
        col_odd = (torch.arange(x.shape[-1], device=x.device) % 2 == 1).view(1, 1, 1, -1)
        out = torch.where(col_odd, yo, ye)
        if self.bias is not None:
            out = out + self.bias.view(1, -1, 1, 1)
        return out

    def forward(self, x):
        """Exactly 7 taps per cell: a 3x1 column convolution on the whole board plus one 2x2 convolution
        per column parity on the other parity's columns (even columns read odd columns at rows r-1, r;
        odd columns read even columns at rows r, r+1)."""
        W = x.shape[-1]
        y = F.conv2d(x, self.kernel0, None, padding=(1, 0))
        xe, xo = x[..., 0::2], x[..., 1::2]
        ne, no = xe.shape[-1], xo.shape[-1]
        # even outputs j (column 2j): odd columns j-1 (left) and j (right), rows r-1 (upper) and r (lower)
        se = F.conv2d(F.pad(xo, (1, ne - no, 1, 0)), self.kernel1, None)
        y[..., 0::2] += se
        if no > 0:
            # odd outputs j (column 2j+1): even columns j (left) and j+1 (right), rows r (upper) and r+1 (lower)
            so = F.conv2d(F.pad(xe, (0, no + 1 - ne, 0, 1)), self.kernel1, None)
            y[..., 1::2] += so
        if self.bias is not None:
            y = y + self.bias.view(1, -1, 1, 1)
        return y


def _conv(cin, cout, hex):
    if hex:
        return HexConv2d(int(cin), int(cout), kernel_size=1, stride=1, bias=False)
    return nn.Conv2d(int(cin), int(cout), kernel_size=3, stride=1, padding="same", bias=False)


class BasicBlock(nn.Module):  # blocks.py:12-41
    def __init__(self, channels, batch_norm=False, hex=True):
        super().__init__()
        layers = [_conv(channels, channels, hex)]
        if batch_norm:
            layers.append(nn.BatchNorm2d(channels))
        layers += [nn.ReLU(), _conv(channels, channels, hex)]
        self.before_shortcut = nn.Sequential(*layers)
        self.shortcut = nn.Sequential()

    def forward(self, x):
        return F.relu(self.before_shortcut(x) + self.shortcut(x))


class Reduce_ValueHead(nn.Module):  # blocks.py:46-92
    def __init__(self, width, num_reduce_layers=4, activation="tanh", batch_norm=False, hex=True):
        super().__init__()
        layers, prev = [], width
        step = (1 - width) / num_reduce_layers
        for layer in range(num_reduce_layers, 0, -1):
            cur = prev + step
            layers.append(_conv(int(prev), int(cur), hex))
            if layer != 1:
                if batch_norm:
                    layers.append(nn.BatchNorm2d(int(cur)))
                layers.append(nn.Tanh() if activation == "tanh" else nn.ReLU())
            prev = cur
        layers += [nn.AdaptiveAvgPool3d(1), nn.Flatten(), nn.Tanh()]
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)


class Reduce_PolicyHead(nn.Module):  # blocks.py:130-170
    def __init__(self, width, policy_channels, num_reduce_layers=2, batch_norm=False, hex=True):
        super().__init__()
        layers, prev = [], width
        step = (policy_channels - width) / num_reduce_layers
        for layer in range(num_reduce_layers, 0, -1):
            cur = prev + step
            layers.append(_conv(int(prev), int(cur), hex))
            if layer != 1:
                if batch_norm:
                    layers.append(nn.BatchNorm2d(int(cur)))
                layers.append(nn.ReLU())
            prev = cur
        self.layers = nn.Sequential(*layers)

    def forward(self, x):
        return self.layers(x)


class RecurrentNet(nn.Module):
    """RecurrentNet.py:18-103 — DeepThinking-style recurrent residual net with recall."""

    def __init__(self, in_channels, policy_channels, num_filters=256, num_blocks=2, recall=True, policy_head="conv",
                 value_head="reduce", value_activation="tanh", hex=True):
        super().__init__()
        self.recurrent = True
        self.recall = recall
        self.num_filters = int(num_filters)
        layers = []
        if recall:
            layers.append(_conv(num_filters + in_channels, num_filters, hex))
        for _ in range(num_blocks):
            layers.append(BasicBlock(self.num_filters, hex=hex))
        self.projection = nn.Sequential(_conv(in_channels, num_filters, hex), nn.ReLU())
        self.recur_module = nn.Sequential(*layers)
        if policy_head != "conv":
            raise ValueError("Unknown choice")
        self.policy_head = Reduce_PolicyHead(num_filters, policy_channels, hex=hex)
        if value_head != "reduce":
            raise ValueError("only the 'reduce' value head is provided")
        self.value_head = Reduce_ValueHead(num_filters, activation=value_activation, hex=hex)

    def forward(self, x, iters_to_do, interim_thought=None, **kwargs):
        thought = self.projection(x) if interim_thought is None else interim_thought
        for _ in range(iters_to_do):
            if self.recall:
                thought = torch.cat([thought, x], 1)
            thought = self.recur_module(thought)
        return (self.policy_head(thought), self.value_head(thought)), thought


class ResNet(nn.Module):  # ResNet.py:14-73
    def __init__(self, in_channels, policy_channels, num_filters=256, num_blocks=4, batch_norm=False, policy_head="conv",
                 value_head="reduce", value_activation="tanh", hex=True):
        super().__init__()
        self.recurrent = False
        layers = [_conv(in_channels, num_filters, hex)]
        if batch_norm:
            layers.append(nn.BatchNorm2d(num_filters))
        layers.append(nn.ReLU())
        self.input_block = nn.Sequential(*layers)
        self.residual_blocks = nn.Sequential(*[BasicBlock(num_filters, batch_norm=batch_norm, hex=hex) for _ in range(num_blocks)])
        self.policy_head = Reduce_PolicyHead(num_filters, policy_channels, batch_norm=batch_norm, hex=hex)
        self.value_head = Reduce_ValueHead(num_filters, activation=value_activation, batch_norm=batch_norm, hex=hex)

    def forward(self, x):
        h = self.residual_blocks(self.input_block(x))
        return self.policy_head(h), self.value_head(h)


class MLP_Network(nn.Module):  # MLP_Network.py:13-76 (what Run.py's preset 0 builds for Tic-Tac-Toe)
    def __init__(self, out_features, hidden_layers=4, neurons_per_layer=64):
        super().__init__()
        self.recurrent = False
        layers = [nn.Flatten(), nn.LazyLinear(64), nn.SiLU()]
        for _ in range(hidden_layers):
            layers += [nn.Linear(neurons_per_layer, neurons_per_layer), nn.SiLU()]
        self.general_module = nn.Sequential(*layers)

        def head(target, act):
            out, prev = [], neurons_per_layer
            step = (target - neurons_per_layer) / 3
            for _ in range(3):
                cur = prev + step
                out += [nn.Linear(int(prev), int(cur)), act()]
                prev = cur
            return nn.Sequential(*out)

        self.policy_head = head(out_features, nn.ReLU)
        self.value_head = head(1, nn.Tanh)

    def forward(self, x):
        h = self.general_module(x)
        return self.policy_head(h), self.value_head(h)


def initialize_parameters(model):
    """Utils/Functions/general_utils.py:8-12: xavier-init exactly the parameters whose name lacks '.weight'."""
    for name, param in model.named_parameters():
        if ".weight" not in name and param.dim() > 1:
            nn.init.xavier_uniform_(param)

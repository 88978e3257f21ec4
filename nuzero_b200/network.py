"""Inference boundary: `Network_Manager` keeps the reference's interface
(Neural_Networks/Network_Manager.py:11-73); `GraphedForward` is the batched bf16 forward, captured
in a CUDA graph, that the search engine's leaf tensor feeds every step."""
import copy

import torch


class Network_Manager:
    def __init__(self, model):
        self.model = model
        self.check_devices()
        if not hasattr(self.model, "recurrent"):
            raise Exception("You need to add a \"recurrent\" bollean atribute to the model,\n"
                            "Specifying if the model is recurrent or not.")
        if not isinstance(self.model.recurrent, bool):
            raise Exception("\"model.recurrent\" must be a bollean atribute specifing if the model is recurrent or not.")

    def is_recurrent(self):
        return self.get_model().recurrent

    def get_model(self):
        return self.model

    def model_to_cpu(self):
        self.model = self.model.to("cpu")

    def model_to_device(self):
        self.model = self.model.to(self.device)

    @staticmethod
    def cuda_is_available():
        return torch.cuda.is_available()

    def check_devices(self):
        self.device = "cuda" if torch.cuda.is_available() else "cpu"
        self.model = self.model.to(self.device)

    def inference(self, state, training, iters_to_do=2, interim_thought=None):
        if not training:
            self.model.eval()
        x = state.to(self.device)
        if not self.model.recurrent:
            if not training:
                with torch.no_grad():
                    return self.model(x)
            return self.model(x)
        if not training:
            with torch.no_grad():
                (p, v), _ = self.model(x, iters_to_do)
            return p, v
        return self.model(x, iters_to_do, interim_thought)


class GraphedForward:
    """Batched forward over the engine's leaf tensor: one CUDA-graph replay per search step.

    The model is cast to `dtype` (bf16 by default — the only tensor-core work of the path) and run on
    engine.leaf; logits are written to engine.policy ([G, A], plane-major like the flat action index
    of Games/Game.py:96-102) and values to engine.value.  Call it right after engine.advance()."""

    def __init__(self, engine, network, iters_to_do=2, dtype=torch.bfloat16, use_graph=True):
        self.e = engine
        model = network.get_model() if hasattr(network, "get_model") else network
        # a private low-precision copy: the caller's Network_Manager keeps its fp32 weights
        self.model = copy.deepcopy(model).to(engine.device).to(dtype).eval()
        self.iters, self.dtype = iters_to_do, dtype
        self.graph = None
        with torch.no_grad():
            if use_graph:
                side = torch.cuda.Stream(engine.device)
                side.wait_stream(torch.cuda.current_stream(engine.device))
                with torch.cuda.stream(side):
                    for _ in range(3):  # warm up cuDNN / cuBLAS heuristics outside the capture
                        self._run()
                torch.cuda.current_stream(engine.device).wait_stream(side)
                self.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph):
                    self._run()

    def _run(self):
        e = self.e
        x = e.leaf if e.leaf.dtype == self.dtype else e.leaf.to(self.dtype)
        if getattr(self.model, "recurrent", False):
            (p, v), _ = self.model(x, self.iters)
        else:
            p, v = self.model(x)
        e.policy.copy_(p.reshape(e.rows, e.A))
        e.value.copy_(v.reshape(e.rows))

    def __call__(self):
        with torch.no_grad():
            if self.graph is not None:
                self.graph.replay()
            else:
                self._run()

"""`graph_loader.GraphDataLoader` -- the module direct_kernel_interface.py:12 imports but the
reference repository does not ship.  Reads the raw int32 CSR files of SURVEY Appendix B
(kernels/graphs/<g>.indptr / .indices) and can also synthesise the BASELINE.json shapes."""
from __future__ import annotations

import glob
import os

import numpy as np
import torch

from maxk_b200 import graphs as _graphs


class GraphDataLoader:
    def __init__(self, base_dir: str = os.path.join("kernels", "graphs")):
        self.base_dir = base_dir

    def get_available_graphs(self):
        return sorted(os.path.splitext(os.path.basename(p))[0]
                      for p in glob.glob(os.path.join(self.base_dir, "*.indptr")))

    def load_graph(self, name: str):
        indptr = np.fromfile(os.path.join(self.base_dir, name + ".indptr"), dtype=np.int32)
        indices = np.fromfile(os.path.join(self.base_dir, name + ".indices"), dtype=np.int32)
        return {"indptr": torch.from_numpy(indptr), "indices": torch.from_numpy(indices),
                "values": torch.ones(indices.shape[0], dtype=torch.float32),
                "v_num": indptr.shape[0] - 1, "e_num": indices.shape[0], "name": name}

    def synthetic(self, shape: str, seed: int = 0, device="cpu", scale: float = 1.0):
        return _graphs.make_graph(shape, seed=seed, device=device, scale=scale)

    @staticmethod
    def to_cuda_tensors(graph_data, device="cuda"):
        out = dict(graph_data)
        for k in ("indptr", "indices", "values"):
            out[k] = graph_data[k].to(device).contiguous()
        return out

    @staticmethod
    def save_graph(graph_data, base_dir: str, name: str):
        os.makedirs(base_dir, exist_ok=True)
        graph_data["indptr"].cpu().numpy().astype(np.int32).tofile(os.path.join(base_dir, name + ".indptr"))
        graph_data["indices"].cpu().numpy().astype(np.int32).tofile(os.path.join(base_dir, name + ".indices"))

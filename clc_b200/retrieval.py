"""Reference retrieval on the GPU: the search the reference runs per sample on the host inside
`LICDataset.__getitem__` (dataloader_ref_cluster.py:149-180) -- a ResNet-50 feature (:41-44, :241-261) and a
ball-tree k-nearest-neighbour query over the dictionary of reference features (:64, :162) -- batched on the
device: `NearestNeighbors` mirrors the part of sklearn's interface the reference uses (`fit`, `kneighbors`), the
distances come from one brute-force pass over the dictionary (`clc_knn_neg_sqdist`) and the k nearest from the
match stage's row-wise top-k kernel."""
import numpy as np
import torch

from ._lib import call, ptr
from .ops import _stream


class NearestNeighbors:
    """sklearn.neighbors.NearestNeighbors(n_neighbors, algorithm=...) as used at dataloader_ref_cluster.py:64:
    Euclidean metric, `fit(X)` then `kneighbors(X) -> (distances, indices)`, nearest first.  `algorithm` is
    accepted and ignored (the result of an exact search does not depend on it).  numpy in -> numpy out
    (float64 distances, int64 indices, like sklearn); CUDA tensors in -> CUDA tensors out (no host sync)."""

    def __init__(self, n_neighbors=5, algorithm="auto", device="cuda"):
        if n_neighbors < 1:
            raise ValueError("Expected n_neighbors > 0. Got %d" % n_neighbors)
        self.n_neighbors, self.algorithm, self.device = n_neighbors, algorithm, torch.device(device)
        self._fit_X = None

    def fit(self, X, y=None):
        X = torch.as_tensor(np.asarray(X) if not isinstance(X, torch.Tensor) else X)
        if X.dim() != 2:
            raise ValueError("Expected 2D array, got %dD array instead" % X.dim())
        self._fit_X = X.to(self.device, torch.float32).contiguous()
        self.n_samples_fit_, self.n_features_in_ = self._fit_X.shape
        return self

    def kneighbors(self, X=None, n_neighbors=None, return_distance=True):
        if self._fit_X is None:
            raise RuntimeError("This NearestNeighbors instance is not fitted yet. Call 'fit' first.")
        k = self.n_neighbors if n_neighbors is None else n_neighbors
        if k > self.n_samples_fit_:
            raise ValueError("Expected n_neighbors <= n_samples_fit, but n_neighbors = %d, n_samples_fit = %d"
                             % (k, self.n_samples_fit_))
        as_numpy = not isinstance(X, torch.Tensor)
        Xq = torch.as_tensor(np.asarray(X)) if as_numpy else X
        if Xq.dim() != 2 or Xq.shape[1] != self.n_features_in_:
            raise ValueError("X has %s features, but NearestNeighbors is expecting %d features as input"
                             % (tuple(Xq.shape), self.n_features_in_))
        Xq = Xq.to(self.device, torch.float32).contiguous()
        Q, N, D = Xq.shape[0], self.n_samples_fit_, self.n_features_in_
        neg = torch.empty(Q, N, dtype=torch.float32, device=self.device)
        val = torch.empty(Q, k, dtype=torch.float32, device=self.device)
        idx = torch.empty(Q, k, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            st = _stream()
            call("clc_knn_neg_sqdist", ptr(Xq), ptr(self._fit_X), Q, N, D, ptr(neg), st)
            call("clc_topk_rows", ptr(neg), Q, N, k, ptr(val), ptr(idx), st)
        dist = torch.sqrt((-val).clamp_min_(0.0))
        if as_numpy:
            d, i = dist.double().cpu().numpy(), idx.long().cpu().numpy()
            return (d, i) if return_distance else i
        return (dist, idx.long()) if return_distance else idx.long()


def make_feature_extractor(state_dict=None, device="cuda"):
    """torchvision ResNet-50 with `fc = Identity` (dataloader_ref_cluster.py:41-44): 2 048-d features.  The
    reference downloads the ImageNet weights (`pretrained=True`); pass them as `state_dict` (no network here)."""
    from torchvision import models
    net = models.resnet50(weights=None)
    if state_dict is not None:
        net.load_state_dict(state_dict)
    net.fc = torch.nn.Identity()
    return net.to(device).eval()


@torch.no_grad()
def retrieve(extractor, searcher, images):
    """Batched version of the per-sample sequence at dataloader_ref_cluster.py:159-163: features of a batch of
    preprocessed images [B, 3, 224, 224] (one ResNet pass), then the n_refs nearest dictionary entries of each.
    -> indices [B, n_refs] (CUDA int64), to be mapped through `feature_to_key`."""
    feats = extractor(images.to(searcher.device, torch.float32))
    return searcher.kneighbors(feats.reshape(feats.shape[0], -1), return_distance=False)

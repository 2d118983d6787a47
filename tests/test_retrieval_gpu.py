"""Reference retrieval on the GPU: clc_b200.retrieval.NearestNeighbors (brute-force scan kernel +
row top-k) against what the reference calls -- sklearn's ball-tree NearestNeighbors(...).kneighbors
(dataloader_ref_cluster.py:64, :162) -- on the same features."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(3000, 2048, 8, 3), (3000, 2048, 21, 5), (37, 12, 11, 1), (500, 64, 3, 32)])
def test_kneighbors_equals_sklearn_ball_tree(shape):
    from sklearn.neighbors import NearestNeighbors as SkNN
    from clc_b200.retrieval import NearestNeighbors
    N, D, Q, k = shape
    rng = np.random.default_rng(sum(shape))
    feats = np.abs(rng.standard_normal((N, D))).astype(np.float32)           # post-ReLU pooled features: >= 0
    queries = (feats[rng.integers(0, N, Q)] + 0.3 * rng.standard_normal((Q, D))).astype(np.float32)
    queries[0] = feats[N // 2]                                              # an exact hit: distance 0
    want_d, want_i = SkNN(n_neighbors=k, algorithm="ball_tree").fit(feats).kneighbors(queries)
    nn = NearestNeighbors(n_neighbors=k, algorithm="ball_tree", device="cuda:0").fit(feats)
    got_d, got_i = nn.kneighbors(queries)
    assert got_i.dtype == np.int64 and got_d.dtype == np.float64
    assert np.array_equal(got_i, want_i)
    assert np.allclose(got_d, want_d, rtol=2e-6, atol=1e-4)                 # fp32 sums vs sklearn's float64
    assert got_i[0, 0] == N // 2 and got_d[0, 0] == 0.0
    # tensors in -> tensors out, no host round trip; indices only
    ti = nn.kneighbors(torch.from_numpy(queries).cuda(), return_distance=False)
    assert ti.is_cuda and np.array_equal(ti.cpu().numpy(), want_i)


def test_kneighbors_argument_errors_follow_sklearn():
    from clc_b200.retrieval import NearestNeighbors
    nn = NearestNeighbors(n_neighbors=3, device="cuda:0")
    with pytest.raises(RuntimeError):
        nn.kneighbors(np.zeros((1, 4), np.float32))
    nn.fit(np.zeros((2, 4), np.float32))
    with pytest.raises(ValueError):
        nn.kneighbors(np.zeros((1, 4), np.float32))          # n_neighbors > n_samples_fit
    with pytest.raises(ValueError):
        nn.kneighbors(np.zeros((1, 5), np.float32), n_neighbors=1)
    with pytest.raises(ValueError):
        NearestNeighbors(n_neighbors=0)


def test_batched_retrieve_pipeline():
    """ResNet-50 (random weights: no network for the ImageNet checkpoint) features of a batch -> nearest dictionary
    entries: the reference's per-sample sequence (extract_feature -> kneighbors, dataloader_ref_cluster.py:159-163)
    as one batched call.  Checked against a brute-force torch search on the same features."""
    from clc_b200 import retrieval
    torch.manual_seed(0)
    ext = retrieval.make_feature_extractor(device="cuda:0")
    imgs = torch.rand(6, 3, 224, 224, device="cuda:0")
    with torch.no_grad():
        dictionary = ext(torch.rand(40, 3, 224, 224, device="cuda:0"))
        feats = ext(imgs)
    assert feats.shape == (6, 2048)
    nn = retrieval.NearestNeighbors(n_neighbors=3, device="cuda:0").fit(dictionary)
    got = retrieval.retrieve(ext, nn, imgs)
    assert got.shape == (6, 3) and got.is_cuda
    d2 = ((dictionary[None].double() - feats[:, None].double()) ** 2).sum(-1)
    want_d = torch.topk(-d2, 3, dim=1).values.neg().sqrt()
    got_d = torch.gather(d2, 1, got).sqrt()
    # (random-weight features of random images are nearly equidistant: compare the distances of the returned
    # neighbours, which is what defines a correct answer under near-ties)
    assert torch.allclose(got_d, want_d, rtol=1e-5, atol=1e-6)

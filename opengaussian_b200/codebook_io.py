"""On-disk codebook format of OpenGaussian: drop-ins for
``train.py::dec2binary / save_kmeans`` (:52-100) and ``utils/opengs_utlis.py::bin2dec / load_code_book``
(:61-88), so that ``render.py`` and the evaluation scripts read what this framework writes and vice versa.

Files (identical names and contents): ``<out_dir>/{root,leaf}_code_book/kmeans_inds.bin`` -- every cluster id
as ``n_bits = ceil(log2(#points))`` bits, most significant bit first, ids concatenated, packed 8 bits per byte
big-endian and zero-padded (what ``bitarray(...).tofile`` writes); ``kmeans_args.npy`` -- a pickled dict
``{'params', 'n_bits', 'total_len'}``; ``kmeans_centers.pth`` -- ``{param: centres}``.

The reference builds a Python list of N * n_bits bools (minutes at 5 M points); here the bits are produced and
packed with tensor ops on the ids' device and only the packed bytes cross to the host.
"""
import math
import os
from collections import OrderedDict

import numpy as np
import torch


def dec2binary(x: torch.Tensor, n_bits=None) -> torch.Tensor:
    """[..., n_bits] bool, most significant bit first (reference train.py:52-61)."""
    if n_bits is None:
        n_bits = torch.ceil(torch.log2(x)).type(torch.int64)
    mask = 2 ** torch.arange(int(n_bits) - 1, -1, -1).to(x.device, x.dtype)
    return x.unsqueeze(-1).bitwise_and(mask).ne(0)


def bin2dec(b: torch.Tensor, bits: int) -> torch.Tensor:
    """Inverse of dec2binary (reference utils/opengs_utlis.py:61-66)."""
    mask = 2 ** torch.arange(bits - 1, -1, -1).to(b.device, torch.int64)
    return torch.sum(mask * b, -1)


def pack_ids(ids: torch.Tensor, n_bits: int) -> np.ndarray:
    """ids [N] -> the bytes of bitarray(dec2binary(ids, n_bits).flatten()).tofile(), as uint8 numpy."""
    ids = ids.reshape(-1).to(torch.int64)
    shifts = torch.arange(n_bits - 1, -1, -1, device=ids.device, dtype=torch.int64)
    out = []
    step = 1 << 20                                    # bound the [chunk, n_bits] temporary
    carry = torch.empty(0, dtype=torch.uint8, device=ids.device)
    weights = torch.tensor([128, 64, 32, 16, 8, 4, 2, 1], dtype=torch.uint8, device=ids.device)
    for lo in range(0, ids.numel(), step):
        bits = ((ids[lo:lo + step].unsqueeze(-1) >> shifts) & 1).to(torch.uint8).reshape(-1)
        bits = torch.cat([carry, bits])
        whole = bits.numel() // 8 * 8
        out.append((bits[:whole].view(-1, 8) * weights).sum(1, dtype=torch.uint8))
        carry = bits[whole:]
    if carry.numel():
        pad = torch.zeros(8 - carry.numel(), dtype=torch.uint8, device=ids.device)
        out.append((torch.cat([carry, pad]).view(-1, 8) * weights).sum(1, dtype=torch.uint8))
    return torch.cat(out).cpu().numpy() if out else np.zeros(0, np.uint8)


def unpack_ids(raw: np.ndarray, n_bits: int, total_len: int, device="cpu") -> torch.Tensor:
    """Inverse of pack_ids: the first total_len bits of `raw` as ids [total_len / n_bits] (int64)."""
    b = torch.from_numpy(np.ascontiguousarray(raw, dtype=np.uint8)).to(device)
    shifts = torch.arange(7, -1, -1, device=b.device, dtype=torch.uint8)
    bits = ((b.unsqueeze(-1) >> shifts) & 1).reshape(-1)[:total_len].to(torch.int64)
    return bin2dec(bits.view(-1, n_bits), n_bits)


def save_kmeans(kmeans_list, quantized_params, out_dir, mode="root"):
    """Reference train.py:63-100.  Same directory layout, file names and bytes."""
    out_dir = os.path.join(out_dir, "root_code_book" if mode == "root" else "leaf_code_book")
    os.makedirs(out_dir, exist_ok=True)
    chunks, total_len, n_bits = [], 0, 0
    pending = np.zeros(0, np.uint8)
    for kmeans in kmeans_list:
        cls_ids = kmeans.cls_ids if mode == "root" else kmeans.leaf_cls_ids
        n_bits = int(np.ceil(np.log2(len(cls_ids))))
        if total_len % 8 == 0:
            chunks.append(pack_ids(cls_ids, n_bits))
        else:   # a previous parameter did not end on a byte boundary: re-pack the joined bit string
            bits = np.unpackbits(np.concatenate(chunks))[:total_len]
            new = dec2binary(cls_ids.reshape(-1).to(torch.int64), n_bits).cpu().numpy().reshape(-1).astype(np.uint8)
            chunks = [np.packbits(np.concatenate([bits, new]))]
        total_len += int(len(cls_ids)) * n_bits
    data = np.concatenate(chunks) if chunks else pending
    with open(os.path.join(out_dir, "kmeans_inds.bin"), "wb") as f:
        f.write(data.tobytes())
    args_dict = {"params": quantized_params, "n_bits": n_bits, "total_len": total_len}
    np.save(os.path.join(out_dir, "kmeans_args.npy"), args_dict)
    centers_dict = {param: (km.centers if mode == "root" else km.leaf_centers)
                    for km, param in zip(kmeans_list, quantized_params)}
    torch.save(centers_dict, os.path.join(out_dir, "kmeans_centers.pth"))


def load_code_book(base_path):
    """Reference utils/opengs_utlis.py:68-88: returns (codebook dict, indices of 'ins_feat' as numpy int64)."""
    codebook = torch.load(os.path.join(base_path, "kmeans_centers.pth"))
    args_dict = np.load(os.path.join(base_path, "kmeans_args.npy"), allow_pickle=True).item()
    quant_params = args_dict["params"]
    raw = np.fromfile(os.path.join(base_path, "kmeans_inds.bin"), dtype=np.uint8)
    indices = unpack_ids(raw, int(args_dict["n_bits"]), int(args_dict["total_len"])).cpu().numpy()
    indices = np.reshape(indices, (len(quant_params), -1))
    indices_dict = OrderedDict()
    for i, key in enumerate(quant_params):
        indices_dict[key] = indices[i]
    return codebook, indices_dict["ins_feat"]


def ids_bits(n_points: int) -> int:
    """Bits per id in kmeans_inds.bin: ceil(log2(#points)) -- the reference sizes them by the number of POINTS."""
    return int(math.ceil(math.log2(n_points)))


CLUSTER_LANG_KEYS = ("leaf_feat", "leaf_score", "occu_count", "leaf_ind")


def save_cluster_lang(model_path, per_leaf_feat, leaf_ave_score, leaf_occu_count, cluster_indices):
    """`cluster_lang.npz` of Stage 3 (reference train.py:951-954): leaf_feat [k1*k2, 512], leaf_score [k1*k2],
    occu_count [k1*k2], leaf_ind [num_pts]; uncompressed `np.savez`, dtypes kept as they are."""
    arrays = [t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
              for t in (per_leaf_feat, leaf_ave_score, leaf_occu_count, cluster_indices)]
    np.savez(os.path.join(model_path, "cluster_lang.npz"), **dict(zip(CLUSTER_LANG_KEYS, arrays)))


def load_cluster_lang(model_path, device="cuda", min_occurrence=5):
    """What render_lerf_by_text.py:56-62 / scripts/eval_scannet.py:135 read back: (leaf_feat, leaf_score,
    occu_count, leaf_ind) on `device`; features of leaves seen fewer than `min_occurrence` times are zeroed (:62)."""
    saved = np.load(os.path.join(model_path, "cluster_lang.npz"))
    feat, score, occu, ind = (torch.from_numpy(saved[k + ".npy"]).to(device) for k in CLUSTER_LANG_KEYS)
    if min_occurrence:
        feat[occu < min_occurrence] *= 0.0
    return feat, score, occu, ind

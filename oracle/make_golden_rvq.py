"""Cross-check oracle/rvq.py against the reference's IN-TREE residual VQ and write a fixture.

Run in the build container only (reads /root/reference):  python -m oracle.make_golden_rvq
`vector_quantize_pytorch.ResidualVQ` (the class the shipped Codec constructs, codec.py:38-49) is a third-party package that
is neither vendored nor installable offline, so oracle/rvq.py cannot be pinned against it.  The reference does carry an
in-tree residual VQ of the same family - `QuarkAudio-HCodec/HCodec-2.0/vq/core_vq.py` `ResidualVectorQuantization`
(EnCodec lineage: nearest code by squared euclidean distance, residual recursion, sum of code vectors; lines 223-238,
394-412).  This script loads seeded codebooks into THAT class, runs its `encode` / `decode` and compares with
`oracle.rvq.rvq_encode` / `rvq_decode`: same indices wherever the fp64 top-2 margin is numerically safe, same de-quantised
vectors.  Output: tests/golden/rvq_intree.npz (inputs, codebooks, the in-tree class's indices and reconstruction) and
tests/golden/rvq_intree_report.json.
"""
import importlib.util
import json
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/QuarkAudio-HCodec/HCodec-2.0/vq/core_vq.py"


def main():
    from oracle import rvq
    spec = importlib.util.spec_from_file_location("ref_core_vq", SRC)
    core = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(core)
    report = {}
    out = {}
    for tag, (dim, K, nq, B, T) in dict(small=(32, 64, 4, 3, 50), h2=(1024, 4096, 16, 2, 40)).items():
        g = torch.Generator().manual_seed(1234 + dim)
        cb = torch.stack([torch.randn(K, dim, generator=g) * 0.35 * 0.85 ** q for q in range(nq)], 0)
        x = torch.randn(B, dim, T, generator=g) * 0.5
        m = core.ResidualVectorQuantization(dim=dim, codebook_size=K, num_quantizers=nq, kmeans_init=False).eval()
        with torch.no_grad():
            for q, layer in enumerate(m.layers):
                layer._codebook.embed.copy_(cb[q])
                layer._codebook.inited.fill_(True)
            ref_idx = m.encode(x)                    # [nq, B, T]
            ref_rec = m.decode(ref_idx)              # [B, dim, T]
        rows = x.transpose(1, 2).reshape(B * T, dim)
        idx, quant = rvq.rvq_encode(rows, cb)
        ref_rows = ref_idx.permute(1, 2, 0).reshape(B * T, nq)
        same = idx == ref_rows
        _, margin = rvq.rvq_margin_audit(rows, cb, idx)
        first_bad = (~same).float().argmax(1)
        unsafe = [float(margin[r, int(first_bad[r])]) for r in range(B * T) if not bool(same[r].all())]
        rec_err = float((rvq.rvq_decode(ref_rows, cb) - ref_rec.transpose(1, 2).reshape(B * T, dim)).abs().max())
        report[tag] = dict(rows=B * T, nq=nq, index_match=float(same.float().mean()), rows_identical=float(same.all(1).float().mean()),
                           max_margin_at_first_difference=max(unsafe) if unsafe else 0.0, dequantise_max_abs_err=rec_err)
        print(tag, report[tag])
        assert rec_err < 1e-5
        assert not unsafe or max(unsafe) < 1e-4, "in-tree RVQ disagrees with the oracle on a numerically safe decision"
        if tag == "small":
            out = dict(x=rows.numpy(), codebooks=cb.numpy(), ref_indices=ref_rows.numpy(),
                       ref_dequant=ref_rec.transpose(1, 2).reshape(B * T, dim).numpy())
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "rvq_intree.npz"), **out)
    json.dump(report, open(os.path.join(ROOT, "tests", "golden", "rvq_intree_report.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

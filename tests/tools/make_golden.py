#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ (run in the build container).

    python tests/tools/make_golden.py

Sources of truth:
* pre/post-processing: the REFERENCE's own functions, imported unmodified from
  /root/reference/catfish/infer.py (oracle/ref_infer.py) - normalize_raw_signal,
  class_from_threshold, correct_short, hp_in_pred;
* network probabilities: the REFERENCE's own op graph, the shipped
  catfish/ResNetRNN/checkpoints/ckpnt-30000.meta, executed node by node by
  tests/tools/meta_graph_interp.py (fp64 and fp32) with the shipped bundle's weights,
  with seeded random weights incl. non-trivial BN statistics, and - for the RNN-only /
  ResNet-only variants - with the graph's own sub-graphs rewired as the reference's
  variants wire them.  Only the H = 16 RNN vector comes from the oracle restatement
  (the graph's consts carry H = 64); its ``source`` entry says so.

The vectors are small (a few hundred KB) so the GPU box, which has no reference
tree, can check both the oracle and the CUDA path against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
from catfish_b200 import synth, weights  # noqa: E402
from oracle import postprocess, ref_infer, tf_graph  # noqa: E402
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import meta_graph_interp as mgi  # noqa: E402
from helpers import random_bn_weights  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def label_patterns(rng):
    pats = [
        [1] * 20 + [0] * 5, [0] * 5 + [1] * 20, [1] * 14 + [0] * 3 + [1] * 15, [1], [0], [1] * 15, [1] * 14,
        [0, 1] * 40, [1] * 100, [0] * 100, [1] * 15 + [0] + [1] * 15, [0] + [1] * 31 + [0], [1] * 32 + [0] * 32 + [1] * 33,
        [2] * 20 + [1] * 3 + [0] * 2 + [3] * 16,
    ]
    for n in (31, 32, 33, 63, 64, 65, 257, 1000):
        p = rng.random() * 0.2 + 0.02
        bits = []
        cur = 0
        while len(bits) < n:
            run = int(rng.geometric(p))
            bits.extend([cur] * run)
            cur ^= 1
        pats.append(bits[:n])
    return pats


def main():
    ref = ref_infer.load()
    rng = np.random.default_rng(2026)
    os.makedirs(OUT, exist_ok=True)

    # ---- post-processing vectors from the reference's own functions
    post = {}
    pats = label_patterns(rng)
    post["n_patterns"] = np.array(len(pats))
    for i, p in enumerate(pats):
        post["pat%d" % i] = np.array(p, np.int64)
        post["pat%d_correct_short" % i] = np.asarray(ref.correct_short(list(p)), np.int64)
        for lab in (1, 0):
            post["pat%d_hp_label%d" % (i, lab)] = np.array(ref.hp_in_pred(list(p), label=lab), np.int64).reshape(-1, 2)
        post["pat%d_hp_ext" % i] = np.array(ref.hp_in_pred(list(p), 3, 0), np.int64).reshape(-1, 2)
    scores = rng.random(500)
    scores[::7] = 0.5
    post["scores"] = scores
    post["scores_labels_0.5"] = np.array(ref.class_from_threshold(scores), np.int64)
    post["scores_labels_0.9"] = np.array(ref.class_from_threshold(scores, 0.9), np.int64)

    # ---- normalisation vectors (reference's normalize_raw_signal)
    raws = [synth.synth_read(n, 7000 + n) for n in (1, 2, 3, 34, 35, 36, 70, 1999, 2000)]
    raws.append(np.array([1, 2, 3, 4], np.int16))
    raws.append(np.array([-32768, 32767, 0, 5, 5, 9], np.int16))          # wide value range
    raws.append(np.array([5, 5, 5, 7, 9, 9], np.int16))
    post["n_raws"] = np.array(len(raws))
    for i, r in enumerate(raws):
        post["raw%d" % i] = r
        with np.errstate(all="ignore"):
            post["raw%d_norm" % i] = np.asarray(ref.normalize_raw_signal(r, "median"), np.float64)
    np.savez_compressed(os.path.join(OUT, "postprocess.npz"), **post)

    # ---- forward vectors: shipped ResNetRNN checkpoint, 4 ragged reads
    w = weights.load_shipped()
    fw = {}
    lengths = [700, 1225, 1999, 2000]        # 1225 = 35 * 35: the "extra window" padding case
    reads = synth.synth_reads(lengths, base_seed=4242)
    torch_graph = tf_graph.TorchGraph(w)
    ref_vars = mgi.load_reference_variables()          # straight from the reference's .index/.data bundle
    fw["n_reads"] = np.array(len(reads))
    fw["source"] = np.array("meta_graph_interp: ckpnt-30000.meta executed op by op")
    for i, r in enumerate(reads):
        norm = ref.normalize_raw_signal(r, "median")
        x, pad = postprocess.pad_and_window(norm)
        p64 = mgi.predictions(x, np.float64, variables=ref_vars)[0][:-pad]
        p32 = mgi.predictions(x, np.float32, variables=ref_vars)[0][:-pad]
        labels = ref.correct_short(ref.class_from_threshold(p32))
        hps = ref.hp_in_pred(labels)
        fw["read%d" % i] = r
        fw["read%d_p64" % i] = p64
        fw["read%d_p32" % i] = p32.astype(np.float32)
        fw["read%d_hps" % i] = np.array(hps, np.int64).reshape(-1, 2)
    # validation heads of the same graph (rnn_class.py:240-241): accuracy/Mean, loss/Mean on read 0
    x, pad = postprocess.pad_and_window(ref.normalize_raw_signal(reads[0], "median"))
    y = (np.random.default_rng(77).random(x.shape) < 0.3).astype(np.float32)
    fw["val_labels"] = y.reshape(-1).astype(np.int8)
    for tag, dt in (("64", np.float64), ("32", np.float32)):
        acc, loss, _ = mgi.accuracy_loss(x, y, dt, variables=ref_vars)
        fw["val_acc" + tag], fw["val_loss" + tag] = np.array(acc), np.array(loss)
    np.savez_compressed(os.path.join(OUT, "forward_resnetrnn_shipped.npz"), **fw)

    # ---- the same graph with seeded random weights and NON-trivial BN statistics (the shipped bundle has
    # moving_mean = 0, moving_variance = 1, which would hide an error in the BN arithmetic)
    wb = random_bn_weights(14)
    x = rng.normal(0, 1.5, size=(24, 35, 1)).astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "forward_resnetrnn_randbn_seed14.npz"), x=x, seed=np.array(14),
                        p64=mgi.predictions(x, np.float64, variables=wb)[0],
                        p32=mgi.predictions(x, np.float32, variables=wb)[0].astype(np.float32),
                        source=np.array("meta_graph_interp"))

    # ---- forward vectors: random-init RNN-only and ResNet-only (A14), 48 windows
    for kind, hpm, seed in (("RNN", dict(layer_size=64, n_layers=3), 11),
                            ("ResNet", dict(layer_size_res=32, n_layers_res=2), 12),
                            ("RNN", dict(layer_size=16, n_layers=2), 13)):
        wr = weights.random_init(kind, seed=seed, **hpm)
        x = rng.normal(0, 1.5, size=(48, 35, 1)).astype(np.float32)
        if hpm.get("layer_size", 64) == 64:
            p64, source = mgi.predictions_variant(kind, x, wr, np.float64), "meta_graph_interp (sub-graphs rewired)"
        else:
            p64, source = tf_graph.forward_np(wr, x, np.float64), "oracle/tf_graph.py (graph consts carry H = 64)"
        tag = "%s_%s" % (kind.lower(), "_".join("%s%d" % (k[0] + k[-1], v) for k, v in sorted(hpm.items())))
        np.savez_compressed(os.path.join(OUT, "forward_%s_seed%d.npz" % (tag, seed)),
                            x=x, p64=p64, seed=np.array(seed), kind=np.array(kind), source=np.array(source),
                            **{"hpm_" + k: np.array(v) for k, v in hpm.items()})
    # ---- chunk merging (next row N1): execute the reference CLI's own source lines (catfish/catfish:58-81, 121-135)
    import copy
    import json
    import textwrap
    src = open("/root/reference/catfish/catfish").read().splitlines()
    body = textwrap.dedent("\n".join(src[57:81]))                 # lines 58-81: the per-file merge block
    center = "\n".join(src[120:135])                              # lines 121-135: def center_hp
    cases = []
    graph_reads = synth.synth_reads([30000, 12000, 4000, 700, 1225], base_seed=600)
    case_inputs = []
    for r in graph_reads:
        hps, n, _ = postprocess.infer_read(r, torch_graph.infer)
        case_inputs.append((hps, n))
    case_inputs += [([], 5000), ([[-11, 46]], 500), ([[-11, 1200]], 3000), ([[100, 160], [150, 210], [2000, 2100]], 2500),
                    ([[5, 40], [30, 1300], [1290, 1400], [4000, 4100]], 4200), ([[2400, 2480]], 2500), ([[0, 10]], 200)]
    for chunk in (1000, 300):
        for hps, n in case_inputs:
            ns = {"hp_positions": copy.deepcopy(hps), "len_read": n, "chunk_size": chunk, "hp_dict": {}, "nonhp_dict": {},
                  "fast5_file": "f"}
            exec(center, ns)
            exec(body, ns)
            cases.append({"hp_positions": hps, "len_read": n, "chunk_size": chunk,
                          "merged": ns["hp_dict"].get("f"), "nonhp": json.loads(json.dumps(ns["nonhp_dict"]["f"]))})
    with open(os.path.join(OUT, "chunks.json"), "w") as f:
        json.dump(cases, f)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()

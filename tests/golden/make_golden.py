"""Generates tests/golden/oracle_checksums.json.

The reference is a Rust crate whose dependencies are not vendored and no cargo exists in the build image, so golden
vectors cannot be produced by running it.  What IS pinned by the reference: the known-answer digests of
src/lib.rs:497-611 (copied as hex below).  The cell checksums are produced by the oracle (oracle/h2sha_oracle.c) after
it passed the mini-MockProver on the same inputs; they freeze the oracle's output so later edits cannot drift silently.
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import mock_prover as MP  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [
    ((128, 128), [b"abc", b""], [0, 0]),                       # test_sha256_correct1
    ((128, 128), [b"\x00", b""], [0, 0]),                      # test_sha256_correct2
    ((128, 128), [b"\x01" * 56, b"\x00\x00\x00"], [0, 0]),     # test_sha256_correct3
    ((128, 128), [bytes(range(192)), bytes(range(64, 256))], [128, 128]),   # shape of test_sha256_correct4 with fixed bytes
    ((64,), [b"\x01" * 55], [0]),                              # BASELINE config 2 shape
    ((320,), [bytes(range(256))], [0]),                        # BASELINE config 4 shape
]

out = {"note": "oracle-generated; digests are the reference's known answers / hashlib", "cases": []}
for sizes, msgs, pre in CASES:
    cfg = O.OracleConfig(max_variable_byte_sizes=sizes)
    reg = O.synthesize(cfg, msgs, pre)
    MP.verify_region(reg, [hashlib.sha256(m).digest() for m in msgs])
    res = O.batch(cfg, reg.layout(), [msgs], [pre])
    out["cases"].append({"max_variable_byte_sizes": list(sizes), "msgs_hex": [m.hex() for m in msgs], "pre_lens": pre,
                         "digests_hex": [d.hex() for d in reg.digests], "checksums": [int(x) for x in res["checksums"][0]],
                         "n_gate": reg.n_gate, "n_lookup": len(reg.lookup_idx), "n_limb": int(reg.dense.shape[0])})
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "oracle_checksums.json"), "w") as f:
    json.dump(out, f, indent=1)
print("wrote", len(out["cases"]), "cases")

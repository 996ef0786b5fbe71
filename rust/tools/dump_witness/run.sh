#!/bin/bash
# One command to pin the oracle's cell placement against the REAL crate (needs cargo + network once; no GPU):
#
#     rust/tools/dump_witness/run.sh            # from the repository root
#
# It builds dump_witness with the reference's own toolchain (nightly-2023-08-12, /root/reference/rust-toolchain; rustup picks it up
# from the rust-toolchain file next to this script), runs the reference's TestCircuit under MockProver for the four input
# pairs of the reference's tests (src/lib.rs:497-611) and compares every advice column of every run with the oracle.
# Expected output, if the recalled halo2-base patterns are right: four lines
#     "8 columns x 131072 rows compared; 0 cells differ"
# Any other outcome names the first differing (column,row) and gate-stream index; DESIGN.md §1 ("Parity status") lists what to suspect first.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
root="$(cd "$here/../../.." && pwd)"
out="${TMPDIR:-/tmp}/h2sha_dump"
mkdir -p "$out"
( cd "$here" && cargo build --release )
bin="$here/target/release/dump_witness"
m192a=$(python3 -c "print(bytes(range(192)).hex())")
m192b=$(python3 -c "print(bytes((i + 64) & 255 for i in range(192)).hex())")
rc=0
run() {   # name hex0 pre0 hex1 pre1
  "$bin" "$out/$1.bin" "$2" "$3" "$4" "$5"
  sha256sum "$out/$1.bin"
  python3 "$root/tools/compare_rust_dump.py" "$out/$1.bin" "$2" "$3" "$4" "$5" || rc=1
}
run correct1 616263 0 "" 0                                            # "abc", ""           (lib.rs:497-527)
run correct2 00 0 "" 0                                                # [0x00], ""          (lib.rs:530-556)
run correct3 "$(python3 -c "print('01' * 56)")" 0 000000 0            # [0x01; 56], [0,0,0] (lib.rs:559-584)
run correct4 "$m192a" 128 "$m192b" 128                                # 192 bytes, precomputed 128 (lib.rs:587-611, fixed bytes instead of thread_rng)
exit $rc

#!/bin/bash
# Pins the oracle to the REAL reference where a Rust toolchain exists; a no-op elsewhere.
#   scripts/gen_reference_goldens.sh [/path/to/bevy_ggrs checkout]
# Builds oracle/ref_harness against the unmodified reference crate, runs it on the seeded populations of the parity
# tests and writes tests/golden/reference_*.json.  tests/test_reference_goldens.py then compares the oracle (CPU) and
# the engine (GPU) with those files checksum for checksum, dt for dt, row for row — and is skipped while they are absent.
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-${BEVY_GGRS_PATH:-/root/reference}}"
if ! command -v cargo >/dev/null 2>&1; then
    echo "cargo not found: the reference cannot be built here (parity stays pinned to the oracle only)"; exit 0
fi
if [ ! -f "$REF/Cargo.toml" ]; then echo "no bevy_ggrs checkout at $REF"; exit 0; fi
cd "$ROOT"
TMP="$(mktemp -d)"
cp -r oracle/ref_harness "$TMP/harness"     # the manifest's path dependency is rewritten to the checkout given
sed -i "s#path = \"../../../reference\"#path = \"$REF\"#" "$TMP/harness/Cargo.toml"
BIN="cargo run --release --quiet --manifest-path $TMP/harness/Cargo.toml --"
run() {  # name entities seed ttl_lo ttl_hi check_distance ticks spawn_rate
    python tests/golden/gen_reference_inputs.py "$TMP/$1.bin" "$2" "$3" "$4" "$5"
    $BIN "$TMP/$1.bin" "$2" "$6" "$7" 60 "$8" > "tests/golden/reference_$1.json"
    echo "wrote tests/golden/reference_$1.json"
}
run small_d4      1000   7    6  40   4  30  0
run despawn_d8    5000   42   3  25   8  40  0
run spawn_d6      300    11   3  40   6  36  40
run c2_100k_d8    100000 0xB200 400 400 8 30 0
rm -rf "$TMP"

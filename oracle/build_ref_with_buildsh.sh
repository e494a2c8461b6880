#!/bin/bash
# Builds the reference library with the reference's OWN build.sh (SURVEY section 8d:
# "REF_BUILDSH = the reference built by its own build.sh"), for the timing baseline of bench.py.
#
#   oracle/build_ref_with_buildsh.sh <reference checkout> <output .so>
#
# build.sh works in the current directory (rm -rf build; cmake; make; ctest) and then runs the reference's full
# benchmark (N up to 1e9: 37 GB, minutes), so: the tree is copied to a scratch directory (the checkout itself is
# read-only and is never written), build.sh runs there unmodified in its own process group, and the group is stopped
# as soon as build.sh announces the benchmark -- the library and the test run are complete by then.  The only thing
# passed in from outside is CC="gcc -fPIC": build.sh's flags (-Ofast -march=native -funroll-loops -ffast-math) are its
# own, and position-independent objects are what lets the static libfabe13.a it produces be wrapped into the shared
# object that bench.py loads.  Nothing from the reference is copied into this repository (the output is git-ignored).
set -u
REF=${1:?reference checkout}
OUT=${2:?output .so}
SCRATCH=$(mktemp -d /tmp/fabe13_ref_buildsh.XXXXXX)
trap 'rm -rf "$SCRATCH"' EXIT
cp -r "$REF"/. "$SCRATCH"/ || exit 1
chmod -R u+w "$SCRATCH"
LOG="$SCRATCH/buildsh.log"
( cd "$SCRATCH" && CC="gcc -fPIC" exec setsid bash ./build.sh ) > "$LOG" 2>&1 &
PID=$!
for _ in $(seq 1 600); do
    if grep -q "Running benchmark" "$LOG" 2>/dev/null; then break; fi
    if ! kill -0 "$PID" 2>/dev/null; then break; fi
    sleep 0.5
done
# stop build.sh and the benchmark it started: the process group of the setsid'ed shell (exactly the processes started above)
PGID=$(ps -o pgid= -p "$PID" 2>/dev/null | tr -d ' ')
if [ -n "${PGID:-}" ]; then kill -- -"$PGID" 2>/dev/null; fi
wait "$PID" 2>/dev/null
if [ ! -f "$SCRATCH/build/libfabe13.a" ]; then
    echo "build.sh did not produce build/libfabe13.a; its log:" >&2
    tail -40 "$LOG" >&2
    exit 1
fi
grep -E "tests passed|Test +#1" "$LOG" | head -3
mkdir -p "$(dirname "$OUT")"
gcc -shared -o "$OUT" -Wl,--whole-archive "$SCRATCH/build/libfabe13.a" -Wl,--no-whole-archive -lm || exit 1
# what build.sh's -march=native meant on this host: the harness loads this build only where every ISA extension exists
grep -m1 '^flags' /proc/cpuinfo | sed 's/.*: //' > "$(dirname "$OUT")/build_host.txt"
echo "built $OUT with the reference's build.sh ($(grep -c . "$LOG") log lines)"

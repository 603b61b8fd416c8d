#!/bin/bash
# A/B harness for kernel tuning: builds libeggsplat variants with -D overrides into eggfusion_b200/variants/
# (git-ignored, travels to the GPU box) -- usage:  profiles/ab_variants.sh build name:"-DX=1 -DY=2" ...
#                                                  profiles/ab_variants.sh run   (on the GPU box; prints stage times)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
V=$ROOT/eggfusion_b200/variants
if [ "$1" = build ]; then
  shift; rm -rf "$V"; mkdir -p "$V"
  for spec in "$@"; do
    name=${spec%%:*}; defs=${spec#*:}
    tmp=$(mktemp -d)
    ( cd "$ROOT/eggfusion_b200/csrc" && for f in *.cu; do
        nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
             --expt-relaxed-constexpr $defs -c "$f" -o "$tmp/${f%.cu}.o" & done; wait
      nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$V/libeggsplat_$name.so" "$tmp"/*.o -cudart shared )
    rm -rf "$tmp"; echo "built $name ($defs)"
  done
else
  for so in "$V"/libeggsplat_*.so; do
    n=$(basename "$so" .so); n=${n#libeggsplat_}
    EGS_LIB=$so python "$ROOT/bench.py" --steps 60 --warmup 5 --no-cpu --no-e2e --no-mapping --no-tracking 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', round(d['ms_per_step'],4), {k: round(v,4) for k,v in d['stage_ms'].items()})"
  done
fi

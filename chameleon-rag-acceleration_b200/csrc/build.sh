#!/bin/bash
# Builds libb200ivfpq.so in-tree for sm_100a (nvcc cross-compiles without a GPU).  Two translation units, compiled in
# parallel: api.cu (C-ABI, orchestration, round-1 kernels) and qlut.cu (the per-query-table filter scan).
set -e
cd "$(dirname "$0")"
OUT=../b200ivfpq/libb200ivfpq.so
OBJ=${B200_BUILD_DIR:-build}
mkdir -p "$OBJ"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v"
pids=()
for tu in api qlut; do
    if [ "$1" != "--incremental" ] || [ ! -f "$OBJ/$tu.o" ] || [ -n "$(find . -maxdepth 1 \( -name '*.cu' -o -name '*.cuh' -o -name '*.h' \) -newer "$OBJ/$tu.o" | head -1)" ]; then
        ( nvcc $FLAGS -c $tu.cu -o "$OBJ/$tu.o" > "$OBJ/$tu.log" 2>&1 || { cat "$OBJ/$tu.log"; exit 1; } ) &
        pids+=($!)
    fi
done
for p in "${pids[@]}"; do wait "$p"; done
nvcc -gencode arch=compute_100a,code=sm_100a --shared -cudart static -o "$OUT" "$OBJ/api.o" "$OBJ/qlut.o"
cat "$OBJ"/api.log "$OBJ"/qlut.log 2>/dev/null | grep -E "spill|error|warning" | sort | uniq -c | sort -rn | head -20
echo "built $OUT"

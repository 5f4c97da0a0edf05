# per-file calls from T host threads vs one chunk_and_hash_batch call (pageable host buffers, like files read by `yams add`)
g++ -std=c++17 -O2 -pthread tools/mt_ingest_bench.cpp -I include -L yams_b200 -lyams_b200 -Wl,-rpath,$PWD/yams_b200 -o /tmp/mt_ingest || exit 1
if [ "$1" != "batch-only" ]; then
echo "| file | threads | calls | avg ms/call | max ms | files/s | GB/s | last timings (thread 0) | |"
for cfg in "65536 1 200" "65536 16 100" "1048576 1 100" "1048576 16 100" "16777216 1 30" "16777216 16 30" "1073741824 1 3"; do timeout 120 /tmp/mt_ingest $cfg; done
fi
echo "batch:"
for cfg in "65536 16 512 batch" "1048576 16 64 batch" "16777216 16 16 batch"; do timeout 120 /tmp/mt_ingest $cfg; done

g++ -std=c++17 -O2 -pthread tools/mt_ingest_bench.cpp -I include -L yams_b200 -lyams_b200 -Wl,-rpath,$PWD/yams_b200 -o /tmp/mt_ingest || exit 1
echo "| file | threads | calls | avg ms/call | max ms | files/s | GB/s | last timings (thread 0) | |"
for pm in 4 64; do echo "POOL_MAX=$pm"; for cfg in "1048576 1 100" "1048576 4 100" "1048576 16 100" "16777216 1 30" "16777216 4 30" "16777216 16 30"; do YAMS_B200_POOL_MAX=$pm timeout 120 /tmp/mt_ingest $cfg; done; done

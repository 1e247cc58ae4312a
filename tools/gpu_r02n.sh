O=gpurun_out/r02n; mkdir -p $O
for ch in 1 2 3 4; do BG_PIPE_CHUNKS=$ch python tools/diag_e2e.py cfg4 100000 0 1 ops 2>&1 | tail -n 1; done
BG_PIPE_CHUNKS=2 BG_PROFILE_HOST=1 python tools/diag_e2e.py cfg4 100000 0 1 ops > $O/prof_cfg4.log 2>&1
g++ -std=c++17 -O2 -pthread -Iinclude tests/cpp/marshal_bench.cpp -Lbiogarden_b200 -lbgalign -Wl,-rpath,$PWD/biogarden_b200 -L/usr/local/cuda/lib64 -lcudart -o tests/cpp/marshal_bench && tests/cpp/marshal_bench 1000000 | tee $O/marshal.json

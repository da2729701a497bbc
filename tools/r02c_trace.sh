cd $GRAFT_REPO_ROOT
O=gpurun_out
export TRACE_LINES=700
python tools/trace_gemm.py 20480 1920 640 > $O/r02c_trace_20480_1920_640.txt 2>&1
python tools/trace_gemm.py 81920 960 320 > $O/r02c_trace_81920_960_320.txt 2>&1
python tools/trace_gemm.py 5120 1280 3840 > $O/r02c_trace_5120_1280_3840.txt 2>&1
python tools/trace_gemm.py 1280 1280 3840 > $O/r02c_trace_1280_1280_3840.txt 2>&1
python tools/trace_gemm.py 81920 320 1280 res > $O/r02c_trace_81920_320_1280_res.txt 2>&1
head -3 $O/r02c_trace_*.txt

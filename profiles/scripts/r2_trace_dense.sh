GVL_LIB_NAME=libgvl_trace2.so python profiles/trace_dense.py cfg2d 2>&1 | tail -10
GVL_LIB_NAME=libgvl_trace2.so python profiles/trace_dense.py cfg3 2>&1 | tail -9
GVL_LIB_NAME=libgvl_trace2.so python profiles/trace_dense.py cfg4 2>&1 | tail -9

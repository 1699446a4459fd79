for d in "" "-DSDFK_X_UNROLL=2" "-DSDFK_X_UNROLL=4" "-DSDFK_X_UNROLL=8"; do echo "== $d"; SDFK_JIT_DEFINES="$d" python tools/time_sample.py 1024 readme | tail -1; done

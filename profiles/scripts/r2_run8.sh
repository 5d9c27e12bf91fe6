for lib in libgvl_b200.so libgvl_mb4.so libgvl_mb6.so; do
  echo "== $lib"
  GVL_LIB_NAME=$lib python profiles/probe_tracks.py 2>&1 | tail -1
  GVL_LIB_NAME=$lib PROBE_VKB=0 python profiles/probe_tracks.py 2>&1 | tail -1
done

# A/B of library builds: standalone class times (tools/probe_classes.py, MBSV_SERIAL_LAUNCHES=1) for every variant given
for v in "$@"; do
  echo "== $v"
  MBSV_B200_LIB=$PWD/multibreak_sv_b200/libmbsv_$v.so MBSV_SERIAL_LAUNCHES=1 python tools/probe_classes.py 2>&1 | tail -14
done

# times the marching kernel for the profiling builds under build_variants/ (GPU box): variants.sh "4 7 8 11"
for v in $1; do
  echo "== B2C_X=$v"; B2C_LIB_PATH=$PWD/build_variants/x$v.so python tools/sweep_rb.py 92 272 2>&1 | grep "1080 64 rb"
done

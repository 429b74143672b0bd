for v in fillq; do
echo "== $v"
ACM_B200_LIB=libacm_b200/_lib/var/$v/libacm_b200.so python tools/dbg_split.py 2>&1 | grep "^stream"
done

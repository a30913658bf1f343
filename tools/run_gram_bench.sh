B=tools/build/gb
for K in 3 5 8 12 16 20; do timeout 60 ${B}_r16s3d0w12b1g1 100003 $K 3 1 | head -1; done
for v in r16s3d0w12b1g1 r16s3d0w12b1g0 r8s4d0w12b1g1 r16s3d4w12b1g1 r16s3d0w8b1g1 r32s2d0w12b1g1; do for N in 1000000 2000000 4000000; do echo "== $v $N"; timeout 60 ${B}_$v $N 20 10 1 | tail -2 | head -1; done; done

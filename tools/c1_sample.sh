# C1 (BASELINE configs[0]): the reference's own perf test at its CI size, `sample 512 25` / `kokkos_sample 512 25`, run through the
# drop-in C++ classes (unmodified caller code), one GPU and all GPUs of the box in one process.
out=gpurun_out/${TAG:-c1}; mkdir -p $out
make -s -C cpp > $out/make.log 2>&1
for prog in sample_b200 kokkos_sample_b200; do
  ( timeout 300 cpp/bin/$prog 512 25 ) > $out/${prog}_512_25.log 2>&1; head -3 $out/${prog}_512_25.log | cut -c1-140
done
( timeout 300 cpp/bin/sample_b200 512 25 pml ) > $out/sample_b200_512_25_pml.log 2>&1; grep "Execution" $out/sample_b200_512_25_pml.log
( timeout 300 cpp/bin/sample_b200 1024 25 ) > $out/sample_b200_1024_25.log 2>&1; head -1 $out/sample_b200_1024_25.log

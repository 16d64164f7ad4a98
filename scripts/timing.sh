mkdir -p gpurun_out
ASM_B200_LIB=tf_face_toolbox_b200/lib/alt/libasoftmax_b200.so ASM_NO_OVERLAP=1 timeout 300 python bench.py --no-cpu-baseline --no-cfg4 --no-graph --steps 2 --warmup 3 2>&1 | grep TIMING | tail -40 > gpurun_out/timing.txt

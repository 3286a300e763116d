cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
export FCP_B200_LIB=$PWD/face_crop_plus_b200/libfcpb200_tl.so
for shape in "3,256,256" "1,256,1024" "3,64,64"; do
FCP_TC_TIMELINE=1 FCP_TL_SHAPE=$shape FCP_TL_SHOT=2 FCP_TL_G0=144 FCP_TL_N=40 timeout 300 python bench.py --batch 16 --det-mb 16 --par-mb 16 --steps 1 --warmup 1 --no-secondary --cpu-sample 0 2>&1 | grep timeline > gpurun_out/tl_$shape.txt
head -50 gpurun_out/tl_$shape.txt
done

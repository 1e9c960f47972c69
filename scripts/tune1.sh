python -m pytest tests/test_stages_gpu.py -q -k "phi_ksi or solve_level" 2>&1 | tail -3
for un in 1 3; do for pf in 1 2; do for vec in 2 4; do
echo "UNROLL=$un PF=$pf VEC=$vec: $(FLOW3D_SWEEP_UNROLL=$un FLOW3D_SWEEP_PF=$pf FLOW3D_SWEEP_VEC=$vec python scripts/run_stage.py sweep --size 512 --reps 5)"
done; done; done
for pf in 0 2; do for vec in 2 4; do
echo "PHIKSI PF=$pf VEC=$vec: $(FLOW3D_PHIKSI_PF=$pf FLOW3D_PHIKSI_VEC=$vec python scripts/run_stage.py phi_ksi --size 512 --reps 5)"
done; done
for sz in 256 128 64; do for vec in 1 2 4; do echo "size $sz VEC=$vec $(FLOW3D_SWEEP_VEC=$vec python scripts/run_stage.py sweep --size $sz --reps 20)"; done; done

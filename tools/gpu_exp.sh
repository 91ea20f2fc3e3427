TAG=${1:-r01n}
O=gpurun_out
mkdir -p $O
( time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pair or slab or mhd_blast_3d or vec_kernels or (vec_equals_generic and 3-32) or implicit_euler or seam or runs_to_t_end or orszag_tang_unsplit or conservation or divB" ) > $O/${TAG}_pytest.log 2>&1
tail -6 $O/${TAG}_pytest.log

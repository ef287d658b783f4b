mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --deselect "tests/test_gpu_big_shapes.py::test_big_shape_slice_against_oracle[C3s-f64]" --deselect "tests/test_gpu_big_shapes.py::test_big_shape_slice_against_oracle[C3s-f32]" --deselect "tests/test_gpu_big_shapes.py::test_big_shape_slice_against_oracle[C4s-f64]" 2>&1 | tail -45 > gpurun_out/r02_fulltests2.txt
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_n1b_C2.json 2> gpurun_out/r02_n1b_C2.err

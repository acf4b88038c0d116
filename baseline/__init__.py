"""Reference arm of the benchmark and of the GPU parity tests: the UNMODIFIED torchode staged under
baseline/_ref/ by scripts/stage_reference.sh (git-ignored; it travels to the GPU box with the gpurun snapshot)."""

#!/bin/bash
bash tools/gpu_env_ab.sh "W2S_LANES=2" 2

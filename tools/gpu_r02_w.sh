#!/bin/bash
bash tools/gpu_ab2.sh "equal" 3

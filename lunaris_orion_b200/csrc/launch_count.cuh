// Process-wide count of kernels launched by this library (bench.py reports it as gpu_launches).
#pragma once
namespace lun {
void note_launch(int n);
}

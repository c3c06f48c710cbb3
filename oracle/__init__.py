"""Test infrastructure: CPU restatement of the reference hot path. Never imported by lunaris_orion_b200/."""

"""TEST INFRASTRUCTURE ONLY: CPU restatement of reference PDLP (see pdlp_cpu_core.h)."""

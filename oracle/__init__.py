"""CPU restatement of the reference algorithm for the hot path — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench cpu_baseline)."""

"""amaranth_twstft_b200: B200-native (sm_100a) post-processing correlator for amaranth_twstft."""

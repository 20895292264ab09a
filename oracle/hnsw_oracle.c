/* placeholder until the HNSW restatement lands */ int cxo_hnsw_placeholder(void){return 0;}

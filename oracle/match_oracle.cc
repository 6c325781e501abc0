extern "C" int orbm_placeholder(){return 0;}

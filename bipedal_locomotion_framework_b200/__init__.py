"""B200-native ContactModels::ContinuousContactModel evaluation (bipedal-locomotion-framework hot
path).  The product is the sm_100a library lib/libblf_ccm.so behind include/blf_ccm.h and the C++17
facade under cpp/; this Python package is the thin harness the tests and bench.py drive it with."""

__all__ = ["synthetic"]

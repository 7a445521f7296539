# A/B plan of the current session (sourced by gpu_session.sh)
one ordered C3 RTCUDA_NO_PIXEL_CULL=1
one ordered C3 X=1
one new C3 X=1
one new C3 RTCUDA_MAX_PATHS=100000000
one new C3 RTCUDA_MAX_PATHS=150000000
one ordered C5s X=1
one new C5s X=1
one new C4 X=1
one new C2 X=1

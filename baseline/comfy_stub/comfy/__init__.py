"""Three-function stand-in for the parts of ComfyUI the reference imports
(``comfy.model_management``; /root/reference/any_device_parallel.py:11, 209, 263, 952).
ComfyUI itself is not installable offline; this stub lets the UNMODIFIED reference
file run as the baseline arm of bench.py and in the differential tests."""

# CUDA 12.9 toolchain is needed at build time: the extension is compiled for sm_100a only.
FROM nvidia/cuda:12.9.0-devel-ubuntu24.04
RUN apt-get update && apt-get install -y --no-install-recommends python3 python3-pip python3-venv ninja-build git && rm -rf /var/lib/apt/lists/*
WORKDIR /workspace/tree-attention-b200
COPY . .
RUN python3 -m venv /opt/venv && /opt/venv/bin/pip install --no-cache-dir -r requirements.txt pytest pytest-timeout hypothesis loguru
ENV PATH=/opt/venv/bin:$PATH
# nvcc cross-compiles without a GPU; the in-tree .so is what the package imports
RUN python -c "import __graft_entry__ as g; g.build()"
CMD ["python3", "model.py"]

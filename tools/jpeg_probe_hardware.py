import sys, json, torch
sys.path.insert(0, ".")
from vision_inspection_system_b200.jpeg import JpegCodec
try:
    JpegCodec(torch.device("cuda", 0), "hardware", True)
    print(json.dumps({"backend": "hardware", "create": "ok"}))
except Exception as e:
    print(json.dumps({"backend": "hardware", "create_error": str(e)}))

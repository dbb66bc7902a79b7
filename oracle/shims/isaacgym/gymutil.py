def parse_device_str(device):
    s = str(device).lower()
    if s == "cpu":
        return "cpu", 0
    parts = s.split(":")
    return parts[0], int(parts[1]) if len(parts) > 1 else 0

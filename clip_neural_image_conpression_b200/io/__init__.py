from .bitstream import MAGIC, VERSION, read_bitstream, read_bitstreams, write_bitstream

__all__ = ["MAGIC", "VERSION", "read_bitstream", "read_bitstreams", "write_bitstream"]

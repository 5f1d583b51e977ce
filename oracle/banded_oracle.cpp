// TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  (placeholder: banded oracle lands with SURVEY 8 rows a10-a15)

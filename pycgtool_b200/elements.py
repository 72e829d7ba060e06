"""Atomic masses by element symbol (upper-cased), amu.

Stands in for the lookup ``mdtraj.element.Element.getBySymbol(sym).mass`` used by
``pycgtool/mapping.py:77-106`` (mdtraj 1.9.7 inherits OpenMM's table).  The
lookup is by *symbol*, so two-letter prefixes of atom names can hit unexpected
elements ("CA" -> calcium, "HO" -> holmium); that quirk is part of the
reference behaviour and is preserved.
"""

MASSES = {
    "H": 1.007947, "D": 2.0135532127, "HE": 4.003, "LI": 6.9412, "BE": 9.0121823, "B": 10.8117,
    "C": 12.01078, "N": 14.00672, "O": 15.99943, "F": 18.99840325, "NE": 20.17976,
    "NA": 22.989769282, "MG": 24.30506, "AL": 26.98153868, "SI": 28.08553, "P": 30.9737622,
    "S": 32.0655, "CL": 35.4532, "AR": 39.9481, "K": 39.09831, "CA": 40.0784, "SC": 44.9559126,
    "TI": 47.8671, "V": 50.94151, "CR": 51.99616, "MN": 54.9380455, "FE": 55.8452,
    "CO": 58.9331955, "NI": 58.69342, "CU": 63.5463, "ZN": 65.4094, "GA": 69.7231, "GE": 72.641,
    "AS": 74.921602, "SE": 78.963, "BR": 79.9041, "KR": 83.7982, "RB": 85.46783, "SR": 87.621,
    "Y": 88.905852, "ZR": 91.2242, "NB": 92.906382, "MO": 95.942, "TC": 98.0, "RU": 101.072,
    "RH": 102.905502, "PD": 106.421, "AG": 107.86822, "CD": 112.4118, "IN": 114.8183,
    "SN": 118.7107, "SB": 121.7601, "TE": 127.603, "I": 126.904473, "XE": 131.2936,
    "CS": 132.90545192, "BA": 137.3277, "LA": 138.905477, "CE": 140.1161, "PR": 140.907652,
    "ND": 144.2423, "PM": 145.0, "SM": 150.362, "EU": 151.9641, "GD": 157.253, "TB": 158.925352,
    "DY": 162.5001, "HO": 164.930322, "ER": 167.2593, "TM": 168.934212, "YB": 173.043,
    "LU": 174.9671, "HF": 178.492, "TA": 180.947882, "W": 183.841, "RE": 186.2071, "OS": 190.233,
    "IR": 192.2173, "PT": 195.0849, "AU": 196.9665694, "HG": 200.592, "TL": 204.38332,
    "PB": 207.21, "BI": 208.980401, "PO": 209.0, "AT": 210.0, "RN": 222.018, "FR": 223.0,
    "RA": 226.0, "AC": 227.0, "TH": 232.038062, "PA": 231.035882, "U": 238.028913, "NP": 237.0,
    "PU": 244.0, "AM": 243.0, "CM": 247.0, "BK": 247.0, "CF": 251.0, "ES": 252.0, "FM": 257.0,
    "MD": 258.0, "NO": 259.0, "LR": 262.0, "RF": 261.0, "DB": 262.0, "SG": 266.0, "BH": 264.0,
    "HS": 269.0, "MT": 268.0, "DS": 281.0, "RG": 272.0, "VS": 0.0,
}


def mass_by_symbol(symbol):
    """Raise KeyError for unknown symbols, like ``Element.getBySymbol``."""
    return MASSES[symbol.strip().upper()]

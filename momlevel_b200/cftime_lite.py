"""A small calendar-aware datetime for time axes when ``cftime`` is not installed.

The reference labels model time with ``cftime`` objects (``test_data/time.py``) and reads the year, the
days in each month and the calendar from them (``util.py:66,79-87,96-102``).  ``cftime`` is not a dependency
here; this class carries exactly the attributes that path reads -- ``year``, ``month``, ``day``,
``daysinmonth``, ``calendar``, ``replace()``, subtraction to a ``datetime.timedelta`` and addition of one --
so that ``steric(dset, annual=True)`` works on a calendar time axis the way the reference call does
(``tests/test_steric.py:158-163``).  Real ``cftime`` objects (what an ``xarray`` Dataset decoded from MOM6
output holds) have the same attributes and take the same code path in ``util.annual_average``.
"""

import datetime as _dt

__all__ = ["Datetime", "month_starts"]

_CUM_365 = (0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334, 365)
_ALIASES = {"365_day": "noleap", "366_day": "all_leap", "standard": "gregorian", "proleptic_gregorian": "gregorian"}


def _canonical(calendar):
    cal = str(calendar).lower()
    cal = _ALIASES.get(cal, cal)
    if cal not in ("noleap", "all_leap", "360_day", "julian", "gregorian"):
        raise ValueError(f"unsupported calendar '{calendar}'")
    return cal


def _is_leap(year, cal):
    if cal == "noleap" or cal == "360_day":
        return False
    if cal == "all_leap":
        return True
    if cal == "julian":
        return year % 4 == 0
    return year % 4 == 0 and (year % 100 != 0 or year % 400 == 0)


def _days_in_month(year, month, cal):
    if cal == "360_day":
        return 30
    n = _CUM_365[month] - _CUM_365[month - 1]
    return n + 1 if (month == 2 and _is_leap(year, cal)) else n


def _days_before_year(year, cal):
    y = year - 1
    if cal == "360_day":
        return 360 * y
    if cal == "noleap":
        return 365 * y
    if cal == "all_leap":
        return 366 * y
    if cal == "julian":
        return 365 * y + y // 4
    return 365 * y + y // 4 - y // 100 + y // 400


def _year_length(year, cal):
    return 360 if cal == "360_day" else (366 if _is_leap(year, cal) else 365)


class Datetime:
    """``Datetime(year, month, day, hour=0, minute=0, second=0, microsecond=0, calendar="noleap")``."""

    __slots__ = ("year", "month", "day", "hour", "minute", "second", "microsecond", "calendar", "_cal")

    def __init__(self, year, month, day, hour=0, minute=0, second=0, microsecond=0, calendar="noleap"):
        self._cal = _canonical(calendar)
        self.calendar = str(calendar)
        if not 1 <= month <= 12 or not 1 <= day <= _days_in_month(year, month, self._cal):
            raise ValueError(f"invalid date {year}-{month}-{day} in calendar {calendar}")
        self.year, self.month, self.day = int(year), int(month), int(day)
        self.hour, self.minute, self.second, self.microsecond = int(hour), int(minute), int(second), int(microsecond)

    @property
    def daysinmonth(self):
        return _days_in_month(self.year, self.month, self._cal)

    def replace(self, **kw):
        f = {k: getattr(self, k) for k in ("year", "month", "day", "hour", "minute", "second", "microsecond", "calendar")}
        f.update(kw)
        return Datetime(**f)

    # ---- arithmetic on a day count since 0001-01-01 of the calendar
    def _ordinal(self):
        if self._cal == "360_day":
            doy = 30 * (self.month - 1) + self.day - 1
        else:
            doy = _CUM_365[self.month - 1] + self.day - 1 + (1 if self.month > 2 and _is_leap(self.year, self._cal) else 0)
        return _days_before_year(self.year, self._cal) + doy

    def _as_delta(self):
        return _dt.timedelta(days=self._ordinal(), hours=self.hour, minutes=self.minute, seconds=self.second,
                             microseconds=self.microsecond)

    @classmethod
    def _from_delta(cls, delta, calendar):
        cal = _canonical(calendar)
        days = delta.days
        year = max(1, days // 366 + 1)
        while _days_before_year(year + 1, cal) <= days:
            year += 1
        doy = days - _days_before_year(year, cal)
        month = 1
        while doy >= _days_in_month(year, month, cal):
            doy -= _days_in_month(year, month, cal)
            month += 1
        rest = delta.seconds
        return cls(year, month, doy + 1, rest // 3600, (rest // 60) % 60, rest % 60, delta.microseconds, calendar=calendar)

    def __sub__(self, other):
        if isinstance(other, Datetime):
            if other._cal != self._cal:
                raise TypeError("cannot subtract dates of different calendars")
            return self._as_delta() - other._as_delta()
        if isinstance(other, _dt.timedelta):
            return Datetime._from_delta(self._as_delta() - other, self.calendar)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, _dt.timedelta):
            return Datetime._from_delta(self._as_delta() + other, self.calendar)
        return NotImplemented

    __radd__ = __add__

    def _key(self):
        return (self.year, self.month, self.day, self.hour, self.minute, self.second, self.microsecond)

    def __eq__(self, other):
        return isinstance(other, Datetime) and self._cal == other._cal and self._key() == other._key()

    def __lt__(self, other):
        return self._key() < other._key()

    def __le__(self, other):
        return self._key() <= other._key()

    def __hash__(self):
        return hash((self._cal,) + self._key())

    def __repr__(self):
        return (f"Datetime({self.year}, {self.month}, {self.day}, {self.hour}, {self.minute}, {self.second}, "
                f"{self.microsecond}, calendar='{self.calendar}')")

    def isoformat(self):
        return f"{self.year:04d}-{self.month:02d}-{self.day:02d}T{self.hour:02d}:{self.minute:02d}:{self.second:02d}"


def month_starts(start_year, nmonths, calendar="noleap"):
    """``nmonths`` consecutive month starts from January of ``start_year`` (``xr.cftime_range(freq="MS")``)."""
    out = []
    for k in range(nmonths):
        out.append(Datetime(start_year + k // 12, k % 12 + 1, 1, calendar=calendar))
    return out

"""Time-quantum views (host side; stays in Go in a real integration).  Mirrors time.go: a time field with quantum "YMDH"
stores every timestamped bit in the standard view and in one view per quantum unit ("standard_2006", "standard_200601",
...; viewsByTime time.go:143-156), and a Row(f=x, from=, to=) query unions the row over the smallest set of views that
covers [from, to) (viewsByTimeRange time.go:158-235).  Go's time.AddDate normalises overflowing days (Oct 31 + 1 month =
Dec 1), which the walk relies on; add_date() reproduces it."""
import datetime as _dt

TIME_FORMAT = "%Y-%m-%dT%H:%M"                     # pilosa.TimeFormat "2006-01-02T15:04"
_UNIT_CHARS = {"Y": 4, "M": 6, "D": 8, "H": 10}


def add_date(t, years=0, months=0, days=0):
    """time.Time.AddDate: field-wise addition, then normalisation"""
    y, m = t.year + years, t.month + months
    y += (m - 1) // 12
    m = (m - 1) % 12 + 1
    return _dt.datetime(y, m, 1, t.hour, t.minute) + _dt.timedelta(days=t.day - 1 + days)


def add_month(t):                                  # addMonth time.go:237-243
    if t.day > 28:
        t = _dt.datetime(t.year, t.month, 1, t.hour)
    return add_date(t, months=1)


def _next_year_gte(t, end):                        # time.go:245-251
    nxt = add_date(t, years=1)
    return nxt.year == end.year or end > nxt


def _next_month_gte(t, end):                       # :253-261
    nxt = add_date(t, months=1)
    return (nxt.year, nxt.month) == (end.year, end.month) or end > nxt


def _next_day_gte(t, end):                         # :263-272
    nxt = add_date(t, days=1)
    return (nxt.year, nxt.month, nxt.day) == (end.year, end.month, end.day) or end > nxt


def view_by_time_unit(name, t, unit):              # :75-88
    return name + "_" + t.strftime({"Y": "%Y", "M": "%Y%m", "D": "%Y%m%d", "H": "%Y%m%d%H"}[unit])


def views_by_time(name, t, quantum):               # :143-156
    full = t.strftime("%Y%m%d%H")
    return [name + "_" + full[:_UNIT_CHARS[u]] for u in quantum if u in _UNIT_CHARS]


def views_by_time_range(name, start, end, quantum):
    """time.go:158-235"""
    has_y, has_m, has_d, has_h = ("Y" in quantum), ("M" in quantum), ("D" in quantum), ("H" in quantum)
    t, out = start, []
    if has_h or has_d or has_m:                    # walk up from the smallest unit to the largest
        while t < end:
            if has_h:
                if not _next_day_gte(t, end):
                    break
                if t.hour != 0:
                    out.append(view_by_time_unit(name, t, "H"))
                    t = t + _dt.timedelta(hours=1)
                    continue
            if has_d:
                if not _next_month_gte(t, end):
                    break
                if t.day != 1:
                    out.append(view_by_time_unit(name, t, "D"))
                    t = add_date(t, days=1)
                    continue
            if has_m:
                if not _next_year_gte(t, end):
                    break
                if t.month != 1:
                    out.append(view_by_time_unit(name, t, "M"))
                    t = add_month(t)
                    continue
            break
    while t < end:                                 # walk back down from the largest unit to the smallest
        if has_y and _next_year_gte(t, end):
            out.append(view_by_time_unit(name, t, "Y"))
            t = add_date(t, years=1)
        elif has_m and _next_month_gte(t, end):
            out.append(view_by_time_unit(name, t, "M"))
            t = add_month(t)
        elif has_d and _next_day_gte(t, end):
            out.append(view_by_time_unit(name, t, "D"))
            t = add_date(t, days=1)
        elif has_h:
            out.append(view_by_time_unit(name, t, "H"))
            t = t + _dt.timedelta(hours=1)
        else:
            break
    return out


def view_time_part(v):                             # :526-534
    last = v.split("_")[-1]
    return last if last.isdigit() else ""


def _lowest_granularity_quantum(views):            # getLowestGranularityQuantum :540-
    lens = {len(view_time_part(v)) for v in views}
    return "".join(u for u in "YMDH" if _UNIT_CHARS[u] in lens)


def min_max_views(views, quantum):                 # :413-468
    views = sorted(views)
    low = _lowest_granularity_quantum(views)
    if low and low in quantum:
        quantum = low
    chars = next((_UNIT_CHARS[u] for u in "YMDH" if u in quantum), 0)
    cand = [v for v in views if len(view_time_part(v)) == chars]
    return (cand[0], cand[-1]) if cand else ("", "")


def time_of_view(v, adj):                          # :474-522
    part = view_time_part(v)
    fmt = {4: "%Y", 6: "%Y%m", 8: "%Y%m%d", 10: "%Y%m%d%H"}.get(len(part))
    if fmt is None:
        raise ValueError(f"invalid time format on view: {v}")
    t = _dt.datetime.strptime(part, fmt)
    if adj:
        t = {4: lambda: add_date(t, years=1), 6: lambda: add_month(t), 8: lambda: add_date(t, days=1), 10: lambda: t + _dt.timedelta(hours=1)}[len(part)]()
    return t


def parse_time(v):
    """parseTime / parsePartialTime time.go:274-411: full "2006-01-02T15:04", a prefix of it down to the year, or epoch seconds"""
    if isinstance(v, bool):
        raise ValueError("arg must be a timestamp")
    if isinstance(v, int):
        return _dt.datetime(1970, 1, 1) + _dt.timedelta(seconds=v)
    if isinstance(v, _dt.datetime):
        return v
    if not isinstance(v, str):
        raise ValueError("arg must be a timestamp")
    for fmt in (TIME_FORMAT, "%Y-%m-%dT%H", "%Y-%m-%d", "%Y-%m", "%Y"):
        try:
            return _dt.datetime.strptime(v, fmt)
        except ValueError:
            pass
    raise ValueError(f"cannot parse time {v!r}")

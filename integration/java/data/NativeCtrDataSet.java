package data;

import com.google.common.collect.Maps;
import nativeps.PsNative;
import org.jblas.FloatMatrix;

import java.io.File;
import java.util.List;
import java.util.Map;

/**
 * Drop-in for `new CTR(new LibsvmParser(), new FileSource(file), batch, thread)` (CTR.java:82-86): same next() /
 * hasNext() / reset() contract as data/DataSet.java:37-67, but lines are read, split and converted by the native reader
 * (ps_reader_*: mmap + producer thread + parallel parsing) instead of String.split / Float.parseFloat per token.
 * Batches the reference would lose to its swallowed exceptions (DataSet.java:96-98) are lost here too.
 * SOURCE ONLY: no JDK in the build image.
 */
public class NativeCtrDataSet extends DataSet {
	private final long reader;
	private final int F = 23, Xn = 45;
	private boolean ended = false;

	public NativeCtrDataSet(File file, int batch, int thread, int offset, int step) {
		super(null, null, batch, 0);               // no Java reader threads
		this.reader = PsNative.readerOpen(file.getPath(), F, Xn, 100000L, batch, offset, step, Math.max(1, thread));
	}
	@Override public void start() {}
	@Override public Map<String, FloatMatrix> next() {
		float[] E = new float[F * batch], X = new float[Xn * batch], W = new float[F * batch], Y = new float[batch];
		int rows = PsNative.readerNext(reader, E, X, W, Y);
		if (rows == 0) { ended = true; return null; }
		Map<String, FloatMatrix> map = Maps.newHashMap();
		map.put("E", wrap(E, F, rows)); map.put("X", wrap(X, Xn, rows)); map.put("W", wrap(W, F, rows)); map.put("Y", wrap(Y, 1, rows));
		return map;
	}
	@Override public boolean hasNext() { return !ended; }
	@Override public void reset() { PsNative.readerReset(reader); ended = false; }
	@Override public Map<String, FloatMatrix> parseFeature(List<List<Feature>> dataList) { throw new UnsupportedOperationException(); }

	private static FloatMatrix wrap(float[] data, int rows, int cols) {
		FloatMatrix m = new FloatMatrix();
		m.data = data; m.rows = rows; m.columns = cols; m.length = rows * cols;   // column-major view, no copy
		return m;
	}
}

/*
 * EtGpu.scala -- the Scala side of the B200 back end of lamp.extratrees.
 *
 * Keeps the four public signatures of extratrees/src/main/scala/lamp/forest/package.scala
 * (buildForestClassification pkg:611-623, buildForestRegression pkg:704-714, predictClassification pkg:542-545,
 * predictRegression pkg:577-580) and the tree ADTs of extratrees.scala:3-63; the bodies call libetgpu.so through
 * the JNI shim jni/etgpu_jni.c.  To switch lamp over, `package.scala`'s four functions delegate to this object
 * (one line each); nothing else in lamp changes.
 *
 * NOT COMPILED IN THIS IMAGE (no JDK / scalac / sbt / saddle): written against Scala 2.13 + saddle-core 4.0.0-M11
 * (build.sbt:83).  The Python ctypes facade lamp_b200/extratrees.py is the same logic, exercised by the tests.
 */
package lamp.extratrees.gpu

import org.saddle._
import lamp.extratrees._

/** One native method per call of include/etgpu.h that the facade needs (jni/etgpu_jni.c). */
object Native {
  System.loadLibrary("etgpu_jni") // links libetgpu.so
  @native def init(devices: Array[Int]): Long
  @native def shutdown(ctx: Long): Unit
  @native def buildClassification(ctx: Long, data: Array[Double], n: Long, d: Int, target: Array[Int],
      weights: Array[Double], numClasses: Int, nMin: Int, k: Int, m: Int, parallelism: Int, bestSplit: Boolean,
      maxDepth: Int, seed: Long): Long
  @native def buildRegression(ctx: Long, data: Array[Double], n: Long, d: Int, target: Array[Double], nMin: Int,
      k: Int, m: Int, parallelism: Int, bestSplit: Boolean, maxDepth: Int, seed: Long): Long
  @native def forestDims(forest: Long, out: Array[Long]): Unit // trees, leaf width, regression flag, total nodes
  @native def exportAll(forest: Long, treeSizes: Array[Int], feature: Array[Int], cut: Array[Double],
      mil: Array[Byte], left: Array[Int], right: Array[Int], leaf: Array[Double]): Unit
  @native def importForest(ctx: Long, leafWidth: Int, regression: Boolean, treeSizes: Array[Int],
      feature: Array[Int], cut: Array[Double], mil: Array[Byte], left: Array[Int], right: Array[Int],
      leaf: Array[Double]): Long
  @native def freeForest(forest: Long): Unit
  @native def predict(ctx: Long, forest: Long, regression: Boolean, samples: Array[Double], n: Long, d: Int,
      out: Array[Double]): Unit
}

object EtGpu {

  /** GPUs to use: ETGPU_DEVICES="0,1,2,3" (several: trees are sharded by tree id inside the library), default "0".
    * The reference hides its cats-effect runtime behind the call (pkg:6,656,675); so does this. */
  private lazy val ctx: Long = {
    val devs = sys.env.getOrElse("ETGPU_DEVICES", "0").split(',').map(_.trim.toInt)
    val h = Native.init(devs)
    sys.addShutdownHook(Native.shutdown(h))
    h
  }

  /** Pre-order arrays of a whole forest as exported by et_forest_export_all. */
  private final case class Flat(treeSizes: Array[Int], feature: Array[Int], cut: Array[Double], mil: Array[Byte],
      left: Array[Int], right: Array[Int], leaf: Array[Double], leafWidth: Int)

  private def export(forest: Long): Flat =
    try {
      val dims = new Array[Long](4)
      Native.forestDims(forest, dims)
      val (m, lw, total) = (dims(0).toInt, dims(1).toInt, dims(3).toInt)
      val f = Flat(new Array[Int](m), new Array[Int](total), new Array[Double](total), new Array[Byte](total),
        new Array[Int](total), new Array[Int](total), new Array[Double](total * lw), lw)
      Native.exportAll(forest, f.treeSizes, f.feature, f.cut, f.mil, f.left, f.right, f.leaf)
      f
    } finally Native.freeForest(forest)

  /** Children have larger pre-order ids than their parent: one backwards pass builds every subtree first. */
  private def decode[T <: AnyRef: scala.reflect.ClassTag](f: Flat)(leaf: Int => T)(
      node: (T, T, Int, Double, Boolean) => T): Seq[T] = {
    var off = 0
    f.treeSizes.toSeq.map { n =>
      val built = new Array[T](n)
      var i = n - 1
      while (i >= 0) {
        val g = off + i
        built(i) =
          if (f.feature(g) < 0) leaf(g)
          else node(built(f.left(g)), built(f.right(g)), f.feature(g), f.cut(g), f.mil(g) != 0)
        i -= 1
      }
      off += n
      built(0)
    }
  }

  def buildForestClassification(
      data: Mat[Double],
      target: Vec[Int],
      sampleWeights: Option[Vec[Double]],
      numClasses: Int,
      nMin: Int,
      k: Int,
      m: Int,
      parallelism: Int,
      bestSplit: Boolean = false,
      maxDepth: Int = Int.MaxValue,
      seed: Long = java.time.Instant.now.toEpochMilli
  ): Seq[ClassificationTree] = {
    // (the library repeats the reference's require(...) checks and reports them as IllegalArgumentException)
    val h = Native.buildClassification(ctx, data.toArray, data.numRows.toLong, data.numCols, target.toArray,
      sampleWeights.map(_.toArray).orNull, numClasses, nMin, k, m, parallelism, bestSplit, maxDepth, seed)
    val f = export(h)
    decode[ClassificationTree](f)(g =>
      ClassificationLeaf(f.leaf.slice(g * f.leafWidth, (g + 1) * f.leafWidth).toSeq))(
      ClassificationNonLeaf(_, _, _, _, _))
  }

  def buildForestRegression(
      data: Mat[Double],
      target: Vec[Double],
      nMin: Int,
      k: Int,
      m: Int,
      parallelism: Int,
      bestSplit: Boolean = false,
      maxDepth: Int = Int.MaxValue,
      seed: Long = java.time.Instant.now.toEpochMilli
  ): Seq[RegressionTree] = {
    val h = Native.buildRegression(ctx, data.toArray, data.numRows.toLong, data.numCols, target.toArray, nMin, k, m,
      parallelism, bestSplit, maxDepth, seed)
    val f = export(h)
    decode[RegressionTree](f)(g => RegressionLeaf(f.leaf(g)))(RegressionNonLeaf(_, _, _, _, _))
  }

  /** Nested case classes -> pre-order arrays (et_forest_import). */
  private final class Flattener(leafWidth: Int) {
    val sizes = Array.newBuilder[Int]
    val feature = Array.newBuilder[Int]; val cut = Array.newBuilder[Double]; val mil = Array.newBuilder[Byte]
    val left = Array.newBuilder[Int]; val right = Array.newBuilder[Int]; val leaf = Array.newBuilder[Double]
    private val zeros = Array.fill(leafWidth)(0d)

    /** Explicit stack (trees are deep): pre-order ids, `right` patched when the left subtree is done. */
    def add[T](root: T)(split: T => Option[(T, T, Int, Double, Boolean)], leafValues: T => Array[Double]): Unit = {
      val f = scala.collection.mutable.ArrayBuffer.empty[Int]; val c = scala.collection.mutable.ArrayBuffer.empty[Double]
      val mi = scala.collection.mutable.ArrayBuffer.empty[Byte]; val l = scala.collection.mutable.ArrayBuffer.empty[Int]
      val r = scala.collection.mutable.ArrayBuffer.empty[Int]; val lv = scala.collection.mutable.ArrayBuffer.empty[Double]
      val stack = scala.collection.mutable.Stack[(T, Int)]((root, -1)) // (subtree, parent waiting for its right id)
      while (stack.nonEmpty) {
        val (t, parent) = stack.pop()
        val id = f.length
        if (parent >= 0) r(parent) = id
        split(t) match {
          case Some((lt, rt, feat, cp, m)) =>
            f += feat; c += cp; mi += (if (m) 1 else 0).toByte; l += id + 1; r += -1; lv ++= zeros
            stack.push((rt, id)); stack.push((lt, -1))
          case None =>
            f += -1; c += Double.NaN; mi += 0.toByte; l += -1; r += -1; lv ++= leafValues(t)
        }
      }
      sizes += f.length; feature ++= f; cut ++= c; mil ++= mi; left ++= l; right ++= r; leaf ++= lv
    }
  }

  def predictClassification(trees: Seq[ClassificationTree], samples: Mat[Double]): Mat[Double] = {
    def width(t: ClassificationTree): Int = t match {
      case ClassificationLeaf(d)               => d.length
      case ClassificationNonLeaf(l, _, _, _, _) => width(l)
    }
    val lw = width(trees.head)
    val fl = new Flattener(lw)
    trees.foreach(t =>
      fl.add[ClassificationTree](t)({
        case ClassificationNonLeaf(l, r, f, c, m) => Some((l, r, f, c, m))
        case _                                    => None
      }, { case ClassificationLeaf(d) => d.toArray; case _ => Array.empty }))
    val h = Native.importForest(ctx, lw, false, fl.sizes.result(), fl.feature.result(), fl.cut.result(),
      fl.mil.result(), fl.left.result(), fl.right.result(), fl.leaf.result())
    try {
      val out = new Array[Double](samples.numRows * lw)
      Native.predict(ctx, h, false, samples.toArray, samples.numRows.toLong, samples.numCols, out)
      Mat(samples.numRows, lw, out) // row-major n x numClasses, like pkg:546-550
    } finally Native.freeForest(h)
  }

  def predictRegression(trees: Seq[RegressionTree], samples: Mat[Double]): Vec[Double] = {
    val fl = new Flattener(1)
    trees.foreach(t =>
      fl.add[RegressionTree](t)({
        case RegressionNonLeaf(l, r, f, c, m) => Some((l, r, f, c, m))
        case _                                => None
      }, { case RegressionLeaf(v) => Array(v); case _ => Array.empty }))
    val h = Native.importForest(ctx, 1, true, fl.sizes.result(), fl.feature.result(), fl.cut.result(),
      fl.mil.result(), fl.left.result(), fl.right.result(), fl.leaf.result())
    try {
      val out = new Array[Double](samples.numRows)
      Native.predict(ctx, h, true, samples.toArray, samples.numRows.toLong, samples.numCols, out)
      Vec(out)
    } finally Native.freeForest(h)
  }
}
